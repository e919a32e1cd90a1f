// huffman.cuh -- Huffman tree description, decode-table construction and stream decode
// (SURVEY.md section 8a rows a8-a11).
//
// replaces: structure/huffman.go:40-107 HuffmanTreeDesc.DecodeFromStream, :112-190 Build,
// :192-264 InitState/DecodeSymbol/DecodeStream, and fse/fse.go:307-390
// DecodeInterleavedFSEStreams (the two-state weight decode).
//
// Decode table: 2^maxBits cells of uint16 = symbol | numberOfBits << 8, in shared memory
// (<= 4 KB at the format's maxBits limit of 11).
#pragma once
#include "fse.cuh"

namespace szb {

constexpr uint32_t kMaxHufBits = 11;
constexpr uint32_t kMaxHufWeights = 255;

// DecodeInterleavedFSEStreams (fse.go:307-390) for the one shape the reference uses: two
// states sharing one table (huffman.go:55-76).  A bad padding yields zero weights because the
// reference discards this function's error (huffman.go:76).  Serial: one lane.
SZB_HD int fse_decode_weights(const uint32_t *table, uint32_t al, const uint8_t *p, uint32_t len, uint8_t *weights,
                              uint32_t *nweights) {
    RevBits r;
    *nweights = 0;
    if (!rev_init(r, p, (int32_t)len) || !rev_skip_padding(r)) return SZB_OK;
    rev_refill(r);
    uint32_t s1 = rev_read(r, al);  // fse.go:329-335 InitState in slice order
    uint32_t s2 = rev_read(r, al);
    uint32_t n = 0;
    for (;;) {  // fse.go:341-388
        uint32_t e = table[s1];
        if (n >= kMaxHufWeights) return SZB_ERR_UNSUPPORTED;  // more weights than byte values: beyond the format (the reference goes on)
        weights[n++] = (uint8_t)fse_code(e);
        rev_refill(r);
        s1 = fse_baseline(e) + rev_read(r, fse_nb(e));
        if (r.remaining < 0) {  // fse.go:362: stream over-read -> flush the other state's symbol
            if (n >= kMaxHufWeights) return SZB_ERR_UNSUPPORTED;  // more weights than byte values: beyond the format (the reference goes on)
            weights[n++] = (uint8_t)fse_code(table[s2]);
            break;
        }
        e = table[s2];
        if (n >= kMaxHufWeights) return SZB_ERR_UNSUPPORTED;  // more weights than byte values: beyond the format (the reference goes on)
        weights[n++] = (uint8_t)fse_code(e);
        rev_refill(r);
        s2 = fse_baseline(e) + rev_read(r, fse_nb(e));
        if (r.remaining < 0) {
            if (n >= kMaxHufWeights) return SZB_ERR_UNSUPPORTED;  // more weights than byte values: beyond the format (the reference goes on)
            weights[n++] = (uint8_t)fse_code(table[s1]);
            break;
        }
    }
    *nweights = n;
    return SZB_OK;
}

// Direct 4-bit weights, high nibble first (huffman.go:86-104).  Serial.
SZB_HD int huf_read_direct_weights(const uint8_t *p, uint32_t avail, uint32_t n, uint8_t *weights) {
    if (((n + 1) >> 1) > avail) return SZB_ERR_UNEXPECTED_EOF;
    for (uint32_t i = 0; i < n; i++) weights[i] = (i & 1) ? (p[i >> 1] & 0xF) : (p[i >> 1] >> 4);
    return SZB_OK;
}

// HuffmanTreeDesc.Build (huffman.go:112-190): weight statistics.  Returns maxBits and the
// per-bit-length symbol counts (rank_count[0..maxBits], the implied last symbol included).
SZB_HD int huf_weight_stats(const uint8_t *weights, uint32_t nw, uint32_t *max_bits_out, uint32_t *last_nb_out,
                            uint32_t *rank_count /* [kMaxHufBits + 2] */) {
    uint64_t sum = 0;
    for (uint32_t i = 0; i < nw; i++) {  // huffman.go:113-120
        uint32_t w = weights[i];
        if (w > 0) sum += (w - 1 < 64) ? (1ull << (w - 1)) : 0;
    }
    uint32_t log = highbit32((uint32_t)sum) + 1;  // huffman.go:125
    uint64_t left = (1ull << log) - sum;
    if (left & (left - 1)) return SZB_ERR_WRONG_SUM_OF_WEIGHTS;  // huffman.go:128-130
    uint32_t last_w = highbit32((uint32_t)left) + 1;            // huffman.go:131
    if (log > kMaxHufBits) return SZB_ERR_UNSUPPORTED;
    for (uint32_t i = 0; i <= kMaxHufBits + 1; i++) rank_count[i] = 0;
    for (uint32_t i = 0; i < nw; i++) {  // huffman.go:138-145
        uint32_t w = weights[i];
        if (w > log + 1) return SZB_ERR_PANIC;  // negative slice index in Go
        rank_count[w ? log + 1 - w : 0]++;
    }
    uint32_t last_nb = log + 1 - last_w;  // huffman.go:147-152
    rank_count[last_nb]++;
    *max_bits_out = log;
    *last_nb_out = last_nb;
    return SZB_OK;
}

// Table fill, serial form (huffman.go:163-187): longest codes first from index 0, symbols
// ascending inside one length.
SZB_HD int huf_build_serial(const uint8_t *weights, uint32_t nw, uint16_t *table, uint32_t *max_bits_out) {
    uint32_t rank_count[kMaxHufBits + 2], rank_idx[kMaxHufBits + 2];
    uint32_t max_bits, last_nb;
    int rc = huf_weight_stats(weights, nw, &max_bits, &last_nb, rank_count);
    if (rc) return rc;
    const uint32_t size = 1u << max_bits;
    rank_idx[max_bits] = 0;
    for (uint32_t i = max_bits; i >= 1; i--) {
        uint32_t nxt = rank_idx[i] + rank_count[i] * (1u << (max_bits - i));
        if (nxt > size) return SZB_ERR_PANIC;
        rank_idx[i - 1] = nxt;
    }
    if (rank_idx[0] != size) return SZB_ERR_CORRUPTED_HUFF_TREE;  // huffman.go:173-175
    for (uint32_t s = 0; s <= nw; s++) {
        uint32_t w = s < nw ? weights[s] : 0;
        uint32_t nb = s < nw ? (w ? max_bits + 1 - w : 0) : last_nb;
        if (nb == 0) continue;
        uint32_t code = rank_idx[nb], l = 1u << (max_bits - nb);
        for (uint32_t j = 0; j < l; j++) table[code + j] = (uint16_t)(s | (nb << 8));
        rank_idx[nb] += l;
    }
    *max_bits_out = max_bits;
    return SZB_OK;
}

// HuffmanDecodingTable.DecodeStream (huffman.go:221-264) for one stream, one lane.
// "state" is the top maxBits bits of the window (zero filled past the stream start), so
// emit Symbols[state], consume NumberOfBits[state] is the same recurrence as
// DecodeSymbol (huffman.go:199-216).  The loop runs while real bits remain and the stream
// must end exactly (ErrDidntUseAllBitsToDecodeHuffman).  expected = the regenerated size of
// this stream; decoding more or fewer symbols is an error (the reference checks streams 1-3,
// literals.go:320,332,349, and panics on the slice bound for the 4th).
SZB_HD int huf_decode_stream(const uint16_t *table, uint32_t max_bits, const uint8_t *p, uint32_t len, uint8_t *out,
                             uint32_t expected) {
    RevBits r;
    if (!rev_init(r, p, (int32_t)len) || !rev_skip_padding(r)) return SZB_ERR_BAD_PADDING;
    uint32_t n = 0;
    while (r.remaining > 0 && n < expected) {
        rev_refill(r);
        uint32_t e = table[rev_peek(r, max_bits)];
        out[n++] = (uint8_t)e;
        rev_skip(r, e >> 8);
    }
    if (r.remaining > 0) return SZB_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH;  // would overflow its slot
    if (r.remaining < 0) return SZB_ERR_DIDNT_USE_ALL_BITS_TO_DECODE_HUFFMAN;
    if (n != expected) return SZB_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH;
    return SZB_OK;
}

}  // namespace szb
