/*
 * szo.c -- ORACLE (test infrastructure, NOT product code).  See szo.h.
 *
 * CPU restatement, in plain C, of the decode path of the Go reference
 * (KillingSpark/sparkzstd, mounted at /root/reference).  Each function names the
 * reference file:line it follows.  Behaviour on VALID frames is bit-exact with the
 * reference; on inputs where the Go code would panic the oracle returns
 * SZO_ERR_PANIC (or the nearest Err* value) instead of crashing.
 *
 * Deliberate divergences (all unreachable from valid frames):
 *   - the window ring buffer (decompression/ringbuffer.go) is replaced by a flat
 *     output buffer; a match that reaches before the start of the frame is an error
 *     (SZO_ERR_CANT_REPEAT_BYTES) instead of reading stale ring contents;
 *   - quirk A.11-9 (overlapping match >= WindowSize after the ring wrapped loses
 *     bytes) is NOT reproduced: the spec-correct bytes are produced.
 */
#include "szo.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
const char *szo_strerror(int code) {
    switch (code) {
    case SZO_OK: return "ok";
    case SZO_ERR_WRONG_MAGICNUMBER: return "ErrWrongMagicnumber";
    case SZO_ERR_CORRUPT_SIZES: return "ErrCorruptSizes";
    case SZO_ERR_OUT_OF_BLOCKS: return "ErrOutOfBlocks";
    case SZO_ERR_ILLEGAL_CONTENT_SIZE_FLAG: return "ErrIllegalContentSizeFlag";
    case SZO_ERR_ILLEGAL_DICTIONARY_ID_FLAG: return "ErrIllegalDictionaryIDFlag";
    case SZO_ERR_NOT_ENOUGH_BYTES_FOR_BLOCK_HEADER: return "ErrNotEnoughBytesForBlockHeader";
    case SZO_ERR_ILLEGAL_BLOCK_TYPE: return "ErrIllegalBlockType";
    case SZO_ERR_ILLEGAL_BLOCK_SIZE: return "ErrIllegalBlockSize";
    case SZO_ERR_WRONG_JUMPTABLE_BYTES: return "ErrWrongJumptableBytes";
    case SZO_ERR_CORRUPTED_JUMPTABLE: return "ErrCorruptedJumptable";
    case SZO_ERR_ILLEGAL_LITERAL_SECTION_TYPE: return "ErrIllegalLiteralSectionType";
    case SZO_ERR_ILLEGAL_LITERAL_SECTION_SIZE_FORMAT: return "ErrIllegalLiteralSectionSizeFormat";
    case SZO_ERR_WRONG_SIZES_BYTES: return "ErrWrongSizesBytes";
    case SZO_ERR_NO_HUFF_TABLE_TO_CARRY_OVER: return "ErrNoHuffTableToCarryOver";
    case SZO_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH: return "ErrStreamDidntDecodeToRightLength";
    case SZO_ERR_WRONG_SUM_OF_WEIGHTS: return "ErrWrongSumOfWeights";
    case SZO_ERR_CORRUPTED_HUFF_TREE: return "ErrCorruptedHuffTree";
    case SZO_ERR_BAD_PADDING: return "ErrBadPadding";
    case SZO_ERR_DIDNT_USE_ALL_BITS_TO_DECODE_HUFFMAN: return "ErrDidntUseAllBitsToDecodeHuffman";
    case SZO_ERR_NOT_ALL_BITS_USED: return "ErrNotAllBitsUsed";
    case SZO_ERR_NO_LL_TABLE_TO_CARRY_OVER: return "ErrNoLLTableToCarryOver";
    case SZO_ERR_NO_ML_TABLE_TO_CARRY_OVER: return "ErrNoMLTableToCarryOver";
    case SZO_ERR_NO_OF_TABLE_TO_CARRY_OVER: return "ErrNoOFTableToCarryOver";
    case SZO_ERR_NOT_ALL_BYTES_USED_WHILE_SEQUENCE_DECODING: return "ErrNotAllBytesUsedWhileSequenceDecoding";
    case SZO_ERR_DIDNT_READ_ALL_PROBABILITIES: return "ErrDidntReadAllProbabilities";
    case SZO_ERR_NO_SYMBOL_FOR_STATE: return "ErrNoSymbolForState";
    case SZO_ERR_CANT_UNWIND: return "ErrCantUnwind";
    case SZO_ERR_DIDNT_COPY_ALL_LITERAL_BYTES: return "ErrDidntCopyAllLiteralBytes";
    case SZO_ERR_IDX_OUT_OF_BOUNDS: return "ErrIdxOutOfBounds";
    case SZO_ERR_CANT_REPEAT_BYTES: return "ErrCantRepeatBytes";
    case SZO_ERR_DIDNT_DUMP_ALL: return "ErrDidntDumpAll";
    case SZO_ERR_UNEXPECTED_EOF: return "io.ErrUnexpectedEOF";
    case SZO_ERR_PANIC: return "reference would panic on this input";
    case SZO_ERR_NOMEM: return "out of memory";
    case SZO_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown error";
    }
}

/* ------------------------------------------------------------------------- */
/* a1: Reversebitstream (bitstream/reversebitstream.go:9-88)                   */

/* NewReversebitstream, reversebitstream.go:9-11 */
void szo_rbits_init(szo_rbits *r, const uint8_t *data, size_t len) {
    r->data = data;
    r->len = (int64_t)len;
    r->offset = (int64_t)len * 8 - 1;
}

/* BitsStillInStream, reversebitstream.go:13-15 */
int64_t szo_rbits_bits_still_in_stream(const szo_rbits *r) { return r->offset; }

/* Read, reversebitstream.go:17-88.  The Go code assembles the value from the partial
 * top byte, the full bytes and the partial lowest byte; the net effect is "bits
 * [offset-n+1 .. offset] of the little-endian integer, zero below bit 0", which is
 * what is computed here. */
uint64_t szo_rbits_read(szo_rbits *r, int n) {
    if (n == 0) return 0; /* :18-20 */
    if (r->offset <= -1) { /* :23-27 reading over the end is allowed */
        r->offset -= n;
        return 0;
    }
    int64_t hi = r->offset;
    int64_t lo = hi - n + 1;
    int64_t l = lo < 0 ? 0 : lo;
    int cnt = (int)(hi - l + 1);
    int64_t first = l >> 3, last = hi >> 3;
    unsigned shift = (unsigned)(l & 7);
    uint64_t v = 0;
    int k = 0;
    for (int64_t b = first; b <= last && k < 8; b++, k++) v |= (uint64_t)r->data[b] << (8 * k);
    v >>= shift;
    if (first + 8 <= last && shift > 0) v |= (uint64_t)r->data[first + 8] << (64 - shift);
    if (cnt < 64) v &= (((uint64_t)1) << cnt) - 1;
    if (lo < 0) v <<= (unsigned)(-lo); /* :67-75 bits beyond offset -1 act as zeros */
    r->offset -= n;
    return v;
}

/* ------------------------------------------------------------------------- */
/* a2: Bitstream (bitstream/bitstream.go:8-90), source = a bounded byte slice   */

/* NewBitstream, bitstream.go:14-16 */
void szo_fbits_init(szo_fbits *b, const uint8_t *data, size_t len) {
    b->data = data;
    b->len = len;
    b->pos = 0;
    b->buffer = 0;
    b->offset = 8;
}

static int fbits_read_byte(szo_fbits *b, uint8_t *out) {
    if (b->pos >= b->len) return SZO_ERR_UNEXPECTED_EOF;
    *out = b->data[b->pos++];
    return SZO_OK;
}

/* UnwindBit, bitstream.go:20-37 */
int szo_fbits_unwind_bit(szo_fbits *b) {
    if (b->offset == 0) return SZO_ERR_CANT_UNWIND;
    b->offset--;
    if (b->offset == 0) {
        if (b->pos == 0) return SZO_ERR_CANT_UNWIND; /* bufio.UnreadByte error */
        b->pos--;                                    /* :29 UnreadByte */
        b->offset = 8;
    }
    return SZO_OK;
}

/* Read, bitstream.go:39-90 */
int szo_fbits_read(szo_fbits *b, int n, uint64_t *out) {
    *out = 0;
    if (n == 0) return SZO_OK;
    uint64_t val = 0;
    if (b->offset + (unsigned)n <= 8) { /* :47-52 */
        uint8_t mask = (uint8_t)((1u << n) - 1);
        val = (uint64_t)((b->buffer >> b->offset) & mask);
        b->offset += (unsigned)n;
        *out = val;
        return SZO_OK;
    }
    unsigned bits_from_buffer = 8 - b->offset; /* :55 */
    if (b->offset < 8) {
        val = (uint64_t)(b->buffer >> b->offset);
        b->offset = 8;
    }
    unsigned remaining = (unsigned)n - bits_from_buffer;
    unsigned from_last = remaining % 8;
    unsigned bytes_needed = remaining / 8;
    for (unsigned i = 0; i < bytes_needed; i++) { /* :67-74 */
        int e = fbits_read_byte(b, &b->buffer);
        if (e) return e;
        val += (uint64_t)b->buffer << (bits_from_buffer + 8 * i);
    }
    if (from_last > 0) { /* :76-84 */
        int e = fbits_read_byte(b, &b->buffer);
        if (e) return e;
        uint8_t mask = (uint8_t)((1u << from_last) - 1);
        val += (uint64_t)(b->buffer & mask) << ((unsigned)n - from_last);
        b->offset = from_last;
    } else {
        b->offset = 8;
    }
    *out = val;
    return SZO_OK;
}

/* ------------------------------------------------------------------------- */
/* a5: BIT_highbit32 (fse/fse.go:235-249) -- De Bruijn variant, highbit(0) == 0 */
uint32_t szo_highbit32(uint32_t value) {
    static const uint32_t debruijn[32] = {0,  9,  1,  10, 13, 21, 2,  29, 11, 14, 16, 18, 22, 25, 3, 30,
                                          8,  12, 20, 28, 15, 17, 24, 7,  19, 27, 23, 6,  26, 5,  4, 31};
    uint32_t v = value;
    v |= v >> 1;
    v |= v >> 2;
    v |= v >> 4;
    v |= v >> 8;
    v |= v >> 16;
    return debruijn[(uint32_t)(v * 0x07C4ACDDu) >> 27];
}

/* ------------------------------------------------------------------------- */
/* a6: predefined distributions and translations (fse/predefined.go:3-78)      */
const int szo_ll_base[36] = {0,  1,  2,  3,  4,  5,  6,  7,  8,    9,     10,    11,
                             12, 13, 14, 15, 16, 18, 20, 22, 24,   28,    32,    40,
                             48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
const int szo_ll_default[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2,
                                2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
const uint8_t szo_ll_extra[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,  0,  0,  1,  1,
                                  1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
const int szo_ml_base[53] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  12,  13,  14,   15,   16,   17,   18,    19,   20,
                             21, 22, 23, 24, 25, 26, 27, 28, 29,  30,  31,  32,   33,   34,   35,   37,    39,   41,
                             43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
const int szo_ml_default[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
const uint8_t szo_ml_extra[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,  0,  0,  0,  0,  0,  0, 0,
                                  0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
const int szo_of_default[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

/* ------------------------------------------------------------------------- */
/* a3, a4, a7: FSE tables (fse/fse.go)                                         */

void szo_fse_table_free(szo_fse_table *t) {
    if (t && t->table) {
        free(t->table);
        t->table = NULL;
    }
}

static int ceil_bytes(int bits) { return bits / 8 + ((bits % 8) != 0); }

/* ReadTabledescriptionFromBitstream, fse.go:28-130 */
int szo_fse_read_table_description(szo_fse_table *t, const uint8_t *src, size_t len, int *bytes_read) {
    szo_fbits bs;
    uint64_t v;
    int e;
    szo_fbits_init(&bs, src, len);
    t->nvalues = 0;
    t->is_rle = 0;
    *bytes_read = 0;
    e = szo_fbits_read(&bs, 4, &v);
    if (e) return e;
    t->accuracy_log = (int)v + 5; /* :37 */
    int64_t remaining = ((int64_t)1) << t->accuracy_log;
    int bits_read = 4;
    int cur = 0;
    while (remaining > 0) { /* :45 */
        uint32_t bits_needed = szo_highbit32((uint32_t)(remaining + 1)) + 1;
        e = szo_fbits_read(&bs, (int)bits_needed, &v);
        bits_read += (int)bits_needed;
        if (e) {
            *bytes_read = ceil_bytes(bits_read);
            return e;
        }
        uint16_t value = (uint16_t)v;
        uint16_t lowermask = (uint16_t)(((uint16_t)1 << (bits_needed - 1)) - 1);                      /* :63 */
        uint16_t thresh = (uint16_t)(((uint16_t)1 << bits_needed) - 1 - (uint16_t)(remaining + 1)); /* :64 */
        if ((uint16_t)(value & lowermask) < thresh) { /* :66-77 "small" number: one bit fewer */
            e = szo_fbits_unwind_bit(&bs);
            if (e) {
                *bytes_read = ceil_bytes(bits_read);
                return e;
            }
            bits_read--;
            value = value & lowermask;
        } else if (value > lowermask) { /* :79-81 */
            value = (uint16_t)(value - thresh);
        }
        if (cur >= SZO_FSE_MAX_SYMBOLS) return SZO_ERR_PANIC;
        t->values[cur++] = (int64_t)value; /* :84 */
        int probability = (int)value - 1;
        if (probability == -1)
            remaining--; /* :89-90 */
        else
            remaining -= probability;
        if (probability == 0) { /* :96-117 zero-run repeat flags */
            uint64_t skip = 3;
            while (skip == 3) {
                e = szo_fbits_read(&bs, 2, &skip);
                bits_read += 2;
                if (e) {
                    *bytes_read = ceil_bytes(bits_read);
                    return e;
                }
                for (uint64_t i = 0; i < skip; i++) {
                    if (cur >= SZO_FSE_MAX_SYMBOLS) return SZO_ERR_PANIC;
                    t->values[cur++] = 1;
                }
            }
        }
    }
    t->nvalues = cur;
    *bytes_read = ceil_bytes(bits_read); /* :120-123 */
    if (remaining != 0) return SZO_ERR_DIDNT_READ_ALL_PROBABILITIES;
    return SZO_OK;
}

/* BuildDecodingTable, fse.go:136-230 */
int szo_fse_build_decoding_table(szo_fse_table *t, const int *symbol_translation, int ntrans,
                                 const uint8_t *extra_bits, int nextra) {
    if (t->accuracy_log > 24) return SZO_ERR_NOMEM;
    int tablesize = 1 << t->accuracy_log;
    int highposition = tablesize - 1;
    int n = t->nvalues;
    int *symbol_next = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    szo_fse_entry *tab = (szo_fse_entry *)calloc((size_t)tablesize, sizeof(szo_fse_entry));
    uint8_t *filled = (uint8_t *)calloc((size_t)tablesize, 1);
    int rc = SZO_OK;
    if (!symbol_next || !tab || !filled) {
        rc = SZO_ERR_NOMEM;
        goto done;
    }
    for (int s = 0; s < n; s++) { /* :146-155 "-1" probabilities go to the top */
        int64_t p = t->values[s] - 1;
        if (p == -1) {
            if (highposition < 0) {
                rc = SZO_ERR_PANIC;
                goto done;
            }
            tab[highposition].symbol = s;
            filled[highposition] = 1;
            highposition--;
            symbol_next[s] = 1;
        } else {
            symbol_next[s] = (int)p;
        }
    }
    int position = 0;
    for (int s = 0; s < n; s++) { /* :160-184 spread */
        int64_t p = t->values[s] - 1;
        if (p > 0) {
            for (int64_t i = 0; i < p; i++) {
                if (filled[position]) { /* :166-169 panic("Overwriting should never happen") */
                    rc = SZO_ERR_PANIC;
                    goto done;
                }
                tab[position].symbol = s;
                filled[position] = 1;
                position += (tablesize >> 1) + (tablesize >> 3) + 3;
                position &= tablesize - 1;
                int guard = 0;
                while (position > highposition) { /* :178-181 skip the low-probability area */
                    position += (tablesize >> 1) + (tablesize >> 3) + 3;
                    position &= tablesize - 1;
                    if (++guard > tablesize) {
                        rc = SZO_ERR_PANIC;
                        goto done;
                    }
                }
            }
        }
    }
    if (position != 0) { /* :186-189 */
        rc = SZO_ERR_PANIC;
        goto done;
    }
    for (int i = 0; i < tablesize; i++) { /* :192-227 */
        if (!filled[i]) { /* nil entry dereference in Go */
            rc = SZO_ERR_PANIC;
            goto done;
        }
        int symbol = tab[i].symbol;
        uint32_t next_state = (uint32_t)symbol_next[symbol];
        symbol_next[symbol]++;
        tab[i].number_of_bits = (uint8_t)((uint32_t)t->accuracy_log - szo_highbit32(next_state));
        tab[i].baseline = (uint16_t)((next_state << tab[i].number_of_bits) - (uint32_t)tablesize);
        if (ntrans > symbol) tab[i].symbol = symbol_translation[symbol]; /* :216-218 */
        if (nextra > symbol) tab[i].additional_bits = extra_bits[symbol]; /* :219-221 */
    }
done:
    free(symbol_next);
    free(filled);
    if (rc != SZO_OK) {
        free(tab);
        return rc;
    }
    if (t->table) free(t->table);
    t->table = tab;
    t->table_size = tablesize;
    return SZO_OK;
}

static int build_default(szo_fse_table *t, const int *dist, int n, int al, const int *trans, int ntrans,
                         const uint8_t *extra, int nextra) {
    memset(t, 0, sizeof(*t));
    t->accuracy_log = al;
    t->nvalues = n;
    for (int i = 0; i < n; i++) t->values[i] = (int64_t)dist[i] + 1; /* value == probability+1 */
    return szo_fse_build_decoding_table(t, trans, ntrans, extra, nextra);
}
/* BuildLiteralLengthsTable, predefined.go:22-30 */
int szo_fse_build_ll_table(szo_fse_table *t) { return build_default(t, szo_ll_default, 36, 6, szo_ll_base, 36, szo_ll_extra, 36); }
/* BuildMatchLengthsTable, predefined.go:52-60 */
int szo_fse_build_ml_table(szo_fse_table *t) { return build_default(t, szo_ml_default, 53, 6, szo_ml_base, 53, szo_ml_extra, 53); }
/* BuildOffsetTable, predefined.go:70-78 */
int szo_fse_build_of_table(szo_fse_table *t) { return build_default(t, szo_of_default, 29, 5, NULL, 0, NULL, 0); }

/* DecodingTable interface (sequences.go:32-40) over FSETable (fse.go:253-301) and
 * RepeatingDecodingTable (sequences.go:42-62) */
static void tbl_init_state(szo_fse_table *t, szo_rbits *src) { /* fse.go:253-257 */
    if (t->is_rle) return;
    t->state = (int64_t)szo_rbits_read(src, t->accuracy_log);
}
static int tbl_peek_symbol(const szo_fse_table *t, int *sym) { /* fse.go:272-278 */
    if (t->is_rle) {
        *sym = t->rle_value;
        return SZO_OK;
    }
    if (t->state > (int64_t)t->table_size) return SZO_ERR_NO_SYMBOL_FOR_STATE;
    if (t->state == (int64_t)t->table_size) return SZO_ERR_PANIC;
    *sym = t->table[t->state].symbol;
    return SZO_OK;
}
static int tbl_additional_bits(const szo_fse_table *t) { /* fse.go:261-263 */
    if (t->is_rle) return t->rle_additional_bits;
    return t->table[t->state].additional_bits;
}
static void tbl_next_state(szo_fse_table *t, szo_rbits *src) { /* fse.go:282-290 */
    if (t->is_rle) return;
    const szo_fse_entry *e = &t->table[t->state];
    uint64_t add = szo_rbits_read(src, e->number_of_bits);
    t->state = (int64_t)e->baseline + (int64_t)add;
}

/* DecodeInterleavedFSEStreams, fse.go:307-390, for the only call shape used
 * (two states sharing one table, huffman.go:55-66).  t2 shares t1->table. */
int szo_fse_decode_interleaved(szo_fse_table *t1, szo_fse_table *t2, const uint8_t *src, size_t len,
                               uint8_t *out, int out_cap, int *nout) {
    szo_rbits bits;
    szo_fse_table *tabs[2] = {t1, t2};
    int bits_read = 0;
    *nout = 0;
    szo_rbits_init(&bits, src, len);
    uint64_t x = 0;
    while (x == 0) { /* :313-320 */
        x = szo_rbits_read(&bits, 1);
        bits_read++;
        if (bits_read > 64) break; /* Go would spin until offset underflows; treat as bad padding */
    }
    if (bits_read > 8) return SZO_ERR_BAD_PADDING; /* :322-324 */
    for (int i = 0; i < 2; i++) tbl_init_state(tabs[i], &bits); /* :329-335 */
    int should_finish = 0;
    while (!should_finish) { /* :341-388 */
        for (int idx = 0; idx < 2; idx++) {
            int sym;
            int e = tbl_peek_symbol(tabs[idx], &sym);
            if (e) return e;
            tbl_next_state(tabs[idx], &bits);
            if (*nout >= out_cap) return SZO_ERR_PANIC;
            out[(*nout)++] = (uint8_t)sym;
            if (szo_rbits_bits_still_in_stream(&bits) < -1) { /* :362 over-read => flush other state */
                int peek_idx = (1 + idx) % 2;
                e = tbl_peek_symbol(tabs[peek_idx], &sym);
                if (e) return e;
                if (*nout >= out_cap) return SZO_ERR_PANIC;
                out[(*nout)++] = (uint8_t)sym;
                should_finish = 1;
                break;
            }
        }
    }
    return SZO_OK;
}

/* ------------------------------------------------------------------------- */
/* a9..a11: Huffman (structure/huffman.go)                                     */

void szo_huf_table_free(szo_huf_table *t) {
    if (!t) return;
    free(t->number_of_bits);
    free(t->symbols);
    t->number_of_bits = NULL;
    t->symbols = NULL;
}

/* HuffmanTreeDesc.DecodeFromStream, huffman.go:40-107.  weights must hold 4096 bytes. */
int szo_huf_decode_tree_desc(const uint8_t *src, size_t len, uint8_t *weights, int *nweights, int *bytes_used) {
    *nweights = 0;
    *bytes_used = 0;
    if (len < 1) return SZO_ERR_UNEXPECTED_EOF;
    uint8_t header = src[0];
    int bytes_read = 1;
    if (header < 128) { /* :48-85 FSE-compressed weights */
        szo_fse_table fset, fset2;
        memset(&fset, 0, sizeof(fset));
        int bs = 0;
        int e = szo_fse_read_table_description(&fset, src + 1, len - 1, &bs);
        bytes_read += bs;
        if (e) {
            *bytes_used = bytes_read;
            return e;
        }
        e = szo_fse_build_decoding_table(&fset, NULL, 0, NULL, 0);
        if (e) {
            *bytes_used = bytes_read;
            return e;
        }
        fset2 = fset; /* :65 shallow copy: separate state, shared table */
        int bitstream_length = (int)header - bs;
        if (bitstream_length < 0) { /* make([]byte, negative) */
            szo_fse_table_free(&fset);
            return SZO_ERR_PANIC;
        }
        if ((size_t)bytes_read + (size_t)bitstream_length > len) {
            szo_fse_table_free(&fset);
            *bytes_used = (int)len;
            return SZO_ERR_UNEXPECTED_EOF;
        }
        /* :76 the error of DecodeInterleavedFSEStreams is DISCARDED by the reference */
        (void)szo_fse_decode_interleaved(&fset, &fset2, src + bytes_read, (size_t)bitstream_length, weights, 4096,
                                         nweights);
        bytes_read += bitstream_length;
        szo_fse_table_free(&fset);
    } else { /* :86-104 direct 4-bit weights, high nibble first */
        int n = (int)header - 127;
        uint8_t buf = 0;
        for (int i = 0; i < n; i++) {
            if (i % 2 == 0) {
                if ((size_t)bytes_read >= len) return SZO_ERR_UNEXPECTED_EOF;
                buf = src[bytes_read++];
                weights[i] = buf >> 4;
            } else {
                weights[i] = buf & 0xF;
            }
        }
        *nweights = n;
    }
    *bytes_used = bytes_read;
    return SZO_OK;
}

/* HuffmanTreeDesc.Build, huffman.go:112-190 */
int szo_huf_build(const uint8_t *weights, int nweights, szo_huf_table *out) {
    memset(out, 0, sizeof(*out));
    uint64_t sum = 0;
    for (int i = 0; i < nweights; i++) { /* :113-120 */
        uint8_t w = weights[i];
        uint64_t weight = 0;
        if (w > 0) weight = (w - 1 < 64) ? ((uint64_t)1 << (w - 1)) : 0;
        sum += weight;
    }
    uint32_t log = szo_highbit32((uint32_t)sum) + 1; /* :125 */
    uint64_t actual_sum = (uint64_t)1 << log;
    uint64_t left_over = actual_sum - sum;
    if ((left_over & (left_over - 1)) != 0) return SZO_ERR_WRONG_SUM_OF_WEIGHTS; /* :128-130 */
    uint32_t last_weight = szo_highbit32((uint32_t)left_over) + 1;              /* :131 */
    int max_bits = (int)log;
    if (max_bits > 20) return SZO_ERR_PANIC; /* Go would allocate 2^maxBits ints; not reachable from valid data */
    int nsym = nweights + 1;
    int *num_bits = (int *)calloc((size_t)nsym, sizeof(int));
    int *rank_count = (int *)calloc((size_t)max_bits + 1, sizeof(int));
    int *rank_idx = (int *)calloc((size_t)max_bits + 1, sizeof(int));
    int size = 1 << max_bits;
    int *tab_bits = (int *)calloc((size_t)size, sizeof(int));
    int *tab_sym = (int *)calloc((size_t)size, sizeof(int));
    int rc = SZO_OK;
    if (!num_bits || !rank_count || !rank_idx || !tab_bits || !tab_sym) {
        rc = SZO_ERR_NOMEM;
        goto fail;
    }
    for (int i = 0; i < nweights; i++) { /* :138-145 */
        int nob = 0;
        if (weights[i] > 0) nob = max_bits + 1 - (int)weights[i];
        if (nob < 0 || nob > max_bits) { /* slice index out of range in Go */
            rc = SZO_ERR_PANIC;
            goto fail;
        }
        num_bits[i] = nob;
        rank_count[nob]++;
    }
    {
        int last_nob = 0; /* :147-152 */
        if (last_weight > 0) last_nob = max_bits + 1 - (int)last_weight;
        if (last_nob < 0 || last_nob > max_bits) {
            rc = SZO_ERR_PANIC;
            goto fail;
        }
        num_bits[nweights] = last_nob;
        rank_count[last_nob]++;
    }
    for (int i = max_bits; i >= 1; i--) { /* :163-171 longest codes at the lowest indices */
        int64_t next = (int64_t)rank_idx[i] + (int64_t)rank_count[i] * ((int64_t)1 << (max_bits - i));
        if (next > size) { /* write past the table */
            rc = SZO_ERR_PANIC;
            goto fail;
        }
        rank_idx[i - 1] = (int)next;
        for (int j = rank_idx[i]; j < rank_idx[i - 1]; j++) tab_bits[j] = i;
    }
    if (rank_idx[0] != size) { /* :173-175 */
        rc = SZO_ERR_CORRUPTED_HUFF_TREE;
        goto fail;
    }
    for (int i = 0; i < nsym; i++) { /* :177-187 within a length, symbols ascending */
        if (num_bits[i] != 0) {
            int code = rank_idx[num_bits[i]];
            int l = 1 << (max_bits - num_bits[i]);
            for (int j = 0; j < l; j++) tab_sym[code + j] = i;
            rank_idx[num_bits[i]] += l;
        }
    }
    out->max_bits = max_bits;
    out->size = size;
    out->number_of_bits = tab_bits;
    out->symbols = tab_sym;
    free(num_bits);
    free(rank_count);
    free(rank_idx);
    return SZO_OK;
fail:
    free(num_bits);
    free(rank_count);
    free(rank_idx);
    free(tab_bits);
    free(tab_sym);
    return rc;
}

/* HuffmanDecodingTable.DecodeStream, huffman.go:221-264 (InitState :192-196,
 * DecodeSymbol :199-216) */
int szo_huf_decode_stream(const szo_huf_table *t, const uint8_t *data, size_t len, uint8_t *out, size_t out_cap,
                          int *nout) {
    szo_rbits bits;
    int bitsum = 0;
    *nout = 0;
    szo_rbits_init(&bits, data, len);
    uint64_t x = 0;
    while (x == 0 && bitsum <= 8) { /* :227-234 */
        x = szo_rbits_read(&bits, 1);
        bitsum++;
    }
    if (bitsum > 8) return SZO_ERR_BAD_PADDING; /* :236-238 */
    int state = (int)szo_rbits_read(&bits, t->max_bits); /* :240 InitState */
    int total = 0;
    uint16_t mask = (t->max_bits >= 16) ? (uint16_t)0xFFFF : (uint16_t)(((uint16_t)1 << t->max_bits) - 1);
    while (szo_rbits_bits_still_in_stream(&bits) + 1 > -(int64_t)t->max_bits) { /* :248 */
        int symbol = t->symbols[state];
        int nb = t->number_of_bits[state];
        uint64_t rest = szo_rbits_read(&bits, nb);
        state = (int)((uint16_t)((state << nb) + (int)rest) & mask); /* :214 */
        if ((size_t)total >= out_cap) return SZO_ERR_PANIC;          /* output[i] out of range */
        out[total++] = (uint8_t)symbol;
    }
    *nout = total;
    if (szo_rbits_bits_still_in_stream(&bits) + 1 != -(int64_t)t->max_bits) /* :257-261 */
        return SZO_ERR_DIDNT_USE_ALL_BITS_TO_DECODE_HUFFMAN;
    return SZO_OK;
}

/* ------------------------------------------------------------------------- */
/* a15: nextOffset (decompression/sequence_execution.go:65-114)                */
int64_t szo_next_offset(int64_t h[3], const szo_sequence *seq) {
    int64_t offset = 0;
    if (seq->offset <= 3 && seq->literal_length > 0) {
        switch (seq->offset) {
        case 1: offset = h[0]; break;
        case 2:
            offset = h[1];
            h[1] = h[0];
            h[0] = offset;
            break;
        case 3:
            offset = h[2];
            h[2] = h[1];
            h[1] = h[0];
            h[0] = offset;
            break;
        default: break; /* offset value 0 cannot be produced by (1<<code)+bits */
        }
    } else if (seq->offset <= 3 && seq->literal_length == 0) {
        switch (seq->offset) {
        case 1:
            offset = h[1];
            h[1] = h[0];
            h[0] = offset;
            break;
        case 2:
            offset = h[2];
            h[2] = h[1];
            h[1] = h[0];
            h[0] = offset;
            break;
        case 3:
            offset = h[0] - 1;
            h[2] = h[1];
            h[1] = h[0];
            h[0] = offset;
            break;
        default: break;
        }
    } else {
        offset = seq->offset - 3;
        h[2] = h[1];
        h[1] = h[0];
        h[0] = offset;
    }
    return offset;
}

/* a17: Ringbuffer.RepeatBeforeIndex/Repeat (ringbuffer.go:197-277) on a flat buffer:
 * append n bytes, byte-serially, from out[pos-oldest ...] (source may overlap dest). */
int szo_match_copy(uint8_t *out, size_t *pos, size_t cap, int64_t n, int64_t oldest) {
    size_t p = *pos;
    if (oldest <= 0 || (uint64_t)oldest > (uint64_t)p) return SZO_ERR_CANT_REPEAT_BYTES; /* ringbuffer.go:203-214 */
    if (n < 0 || p + (size_t)n > cap) return SZO_ERR_PANIC;
    const uint8_t *s = out + p - (size_t)oldest;
    uint8_t *d = out + p;
    if (oldest >= n) {
        memcpy(d, s, (size_t)n); /* ringbuffer.go:197-233 non-overlapping copy via repeatBuf */
    } else {
        for (int64_t i = 0; i < n; i++) d[i] = s[i]; /* ringbuffer.go:262-273 */
    }
    *pos = p + (size_t)n;
    return SZO_OK;
}

/* ------------------------------------------------------------------------- */
/* growable output                                                            */
typedef struct {
    uint8_t *buf;
    size_t len, cap;
} outbuf;

static int out_reserve(outbuf *o, size_t extra) {
    if (o->len + extra <= o->cap) return SZO_OK;
    size_t ncap = o->cap ? o->cap : 4096;
    while (ncap < o->len + extra) ncap *= 2;
    uint8_t *nb = (uint8_t *)realloc(o->buf, ncap);
    if (!nb) return SZO_ERR_NOMEM;
    o->buf = nb;
    o->cap = ncap;
    return SZO_OK;
}

/* ------------------------------------------------------------------------- */
/* trace helpers                                                              */
void szo_trace_free(szo_trace *t) {
    if (!t) return;
    free(t->blocks);
    free(t->literals);
    free(t->sequences);
    free(t->real_offsets);
    memset(t, 0, sizeof(*t));
}

static szo_block_trace *trace_new_block(szo_trace *t) {
    if (t->nblocks == t->blocks_cap) {
        size_t nc = t->blocks_cap ? t->blocks_cap * 2 : 16;
        szo_block_trace *nb = (szo_block_trace *)realloc(t->blocks, nc * sizeof(*nb));
        if (!nb) return NULL;
        t->blocks = nb;
        t->blocks_cap = nc;
    }
    szo_block_trace *b = &t->blocks[t->nblocks++];
    memset(b, 0, sizeof(*b));
    return b;
}
static int trace_add_literals(szo_trace *t, const uint8_t *p, size_t n, int rle) {
    if (t->nliterals + n > t->literals_cap) {
        size_t nc = t->literals_cap ? t->literals_cap : 4096;
        while (nc < t->nliterals + n) nc *= 2;
        uint8_t *nb = (uint8_t *)realloc(t->literals, nc);
        if (!nb) return SZO_ERR_NOMEM;
        t->literals = nb;
        t->literals_cap = nc;
    }
    if (rle)
        memset(t->literals + t->nliterals, p[0], n);
    else
        memcpy(t->literals + t->nliterals, p, n);
    t->nliterals += n;
    return SZO_OK;
}
static int trace_reserve_seq(szo_trace *t, size_t n) {
    if (t->nsequences + n > t->sequences_cap) {
        size_t nc = t->sequences_cap ? t->sequences_cap : 1024;
        while (nc < t->nsequences + n) nc *= 2;
        szo_sequence *ns = (szo_sequence *)realloc(t->sequences, nc * sizeof(*ns));
        if (!ns) return SZO_ERR_NOMEM;
        t->sequences = ns;
        int64_t *no = (int64_t *)realloc(t->real_offsets, nc * sizeof(*no));
        if (!no) return SZO_ERR_NOMEM;
        t->real_offsets = no;
        t->sequences_cap = nc;
    }
    return SZO_OK;
}

/* ------------------------------------------------------------------------- */
/* per-frame decoder state = FrameDecompressor (framedecompressor.go:14-40)    */
typedef struct {
    /* tables remembered across blocks: PreviousBlock.* (framedecompressor.go:283-294) */
    szo_fse_table *prev_ll, *prev_of, *prev_ml;
    szo_huf_table *prev_huf;
    /* every table allocated for this frame (freed at the end) */
    void **owned;
    int *owned_kind; /* 0 fse, 1 huf */
    size_t nowned, owned_cap;
    int64_t offset_history[3];
    /* scratch, mirrors literalsDataBuf / sequences (framedecompressor.go:26-31) */
    uint8_t *lit_data;
    szo_sequence *seqs;
    size_t seqs_cap;
} frame_state;

static int own(frame_state *fs, void *p, int kind) {
    if (fs->nowned == fs->owned_cap) {
        size_t nc = fs->owned_cap ? fs->owned_cap * 2 : 16;
        void **no = (void **)realloc(fs->owned, nc * sizeof(void *));
        if (!no) return SZO_ERR_NOMEM;
        fs->owned = no;
        int *nk = (int *)realloc(fs->owned_kind, nc * sizeof(int));
        if (!nk) return SZO_ERR_NOMEM;
        fs->owned_kind = nk;
        fs->owned_cap = nc;
    }
    fs->owned[fs->nowned] = p;
    fs->owned_kind[fs->nowned] = kind;
    fs->nowned++;
    return SZO_OK;
}
static szo_fse_table *new_fse(frame_state *fs) {
    szo_fse_table *t = (szo_fse_table *)calloc(1, sizeof(*t));
    if (!t) return NULL;
    if (own(fs, t, 0)) {
        free(t);
        return NULL;
    }
    return t;
}
static void frame_state_free(frame_state *fs) {
    for (size_t i = 0; i < fs->nowned; i++) {
        if (fs->owned_kind[i] == 0) {
            szo_fse_table_free((szo_fse_table *)fs->owned[i]);
        } else {
            szo_huf_table_free((szo_huf_table *)fs->owned[i]);
        }
        free(fs->owned[i]);
    }
    free(fs->owned);
    free(fs->owned_kind);
    free(fs->lit_data);
    free(fs->seqs);
}

#define MAX_BLOCK (128 * 1024)

/* Literal section state handed from decode to execution (literals.go:10-19) */
typedef struct {
    int type, streams;
    int regen;
    const uint8_t *data; /* Data; for RLE only data[0] is meaningful */
    int data_len;        /* len(ls.Data) */
    int data_read;       /* ls.dataRead */
    int total_bytes;     /* header + tree + payload bytes consumed from the block */
    int huf_max_bits;
} lit_section;

/* LiteralSection.DecodeNextLiteralsSection, literals.go:209-373 (+ header helpers :46-204) */
static int decode_literals_section(frame_state *fs, const uint8_t *src, size_t len, lit_section *ls) {
    size_t pos = 0;
    uint8_t hb[6] = {0, 0, 0, 0, 0, 0};
    memset(ls, 0, sizeof(*ls));
    if (len < 1) return SZO_ERR_UNEXPECTED_EOF;
    hb[0] = src[pos++];
    int header_bytes = 1;
    ls->type = hb[0] & 3; /* :67-81 DecodeType */
    int sizeformat = (hb[0] >> 2) & 3;
    int needed; /* :162-204 BytesNeededToDecodeSizes */
    if (ls->type == 0 || ls->type == 1)
        needed = (sizeformat == 1) ? 2 : (sizeformat == 3) ? 3 : 1;
    else
        needed = (sizeformat == 2) ? 4 : (sizeformat == 3) ? 5 : 3;
    if (needed > 1) {
        if (len < (size_t)needed) return SZO_ERR_UNEXPECTED_EOF;
        for (int i = 1; i < needed; i++) hb[i] = src[pos++];
        header_bytes += needed - 1;
    }
    int regen = 0, comp = 0;
    if (ls->type == 0 || ls->type == 1) { /* :95-124 DecodeSizes raw/rle */
        ls->streams = 1;
        switch (sizeformat) {
        case 0:
        case 2: regen = hb[0] >> 3; break;
        case 1: regen = (hb[0] >> 4) + ((int)hb[1] << 4); break;
        default: regen = (hb[0] >> 4) + (int)(((uint32_t)hb[1] << 4) + ((uint32_t)hb[2] << 12)); break;
        }
        comp = (ls->type == 0) ? regen : 1;
    } else { /* :125-157 compressed / treeless */
        ls->streams = 4;
        uint32_t sizes = ((uint32_t)hb[0] | ((uint32_t)hb[1] << 8) | ((uint32_t)hb[2] << 16) |
                          ((uint32_t)(needed > 3 ? hb[3] : 0) << 24)) >> 4;
        switch (sizeformat) {
        case 0:
            ls->streams = 1; /* fallthrough */
        case 1:
            regen = (int)(sizes & 0x3FF);
            comp = (int)((sizes >> 10) & 0x3FF);
            break;
        case 2:
            regen = (int)(sizes & 0x3FFF);
            comp = (int)((sizes >> 14) & 0x3FFF);
            break;
        default:
            regen = (int)(sizes & 0x3FFFF);
            comp = (int)(((sizes >> 18) & 0x3FFFF) + ((uint32_t)hb[4] << 10));
            break;
        }
    }
    ls->regen = regen;
    szo_huf_table *table = NULL;
    if (ls->type == 3) { /* :247-252 treeless: carry over */
        table = fs->prev_huf;
        if (!table) return SZO_ERR_NO_HUFF_TABLE_TO_CARRY_OVER;
    }
    int tree_bytes = 0;
    if (ls->type == 2) { /* :254-267 */
        uint8_t weights[4096];
        int nweights = 0;
        int e = szo_huf_decode_tree_desc(src + pos, len - pos, weights, &nweights, &tree_bytes);
        if (e) return e;
        pos += (size_t)tree_bytes;
        table = (szo_huf_table *)calloc(1, sizeof(*table));
        if (!table) return SZO_ERR_NOMEM;
        if (own(fs, table, 1)) {
            free(table);
            return SZO_ERR_NOMEM;
        }
        e = szo_huf_build(weights, nweights, table);
        if (e) return e;
        comp -= tree_bytes;
    }
    uint16_t s1 = 0, s2 = 0, s3 = 0;
    if (ls->streams == 4) { /* :270-279 jump table (its validation error is discarded, :277) */
        if (len - pos < 6) return SZO_ERR_UNEXPECTED_EOF;
        s1 = (uint16_t)(src[pos] | (src[pos + 1] << 8));
        s2 = (uint16_t)(src[pos + 2] | (src[pos + 3] << 8));
        s3 = (uint16_t)(src[pos + 4] | (src[pos + 5] << 8));
        pos += 6;
        header_bytes += 6;
        comp -= 6;
    }
    if (comp < 0 || comp > MAX_BLOCK) return SZO_ERR_PANIC; /* :283 slice bounds */
    if (len - pos < (size_t)comp) return SZO_ERR_UNEXPECTED_EOF;
    const uint8_t *cdata = src + pos;
    pos += (size_t)comp;
    if (ls->type == 0 || ls->type == 1) { /* :290-292 */
        ls->data = cdata;
        ls->data_len = comp;
    } else { /* :295-371 */
        if (regen > MAX_BLOCK) return SZO_ERR_PANIC;
        uint8_t *output = fs->lit_data;
        ls->data = output;
        ls->data_len = regen;
        ls->huf_max_bits = table->max_bits;
        if (ls->streams == 1) { /* :299-304; the decoded count is NOT compared with regen */
            int n;
            int e = szo_huf_decode_stream(table, cdata, (size_t)comp, output, (size_t)regen, &n);
            if (e) return e;
        } else {
            int normal = (regen + 3) / 4; /* :306-311 */
            int last = regen - 3 * normal;
            if (last < 0) return SZO_ERR_PANIC;
            int low = 0, high = (int)s1, n1, n2, n3, n4, e;
            if (high > comp) return SZO_ERR_CORRUPTED_JUMPTABLE; /* Go re-slices into stale capacity; corrupt only */
            e = szo_huf_decode_stream(table, cdata + low, (size_t)(high - low), output, (size_t)normal, &n1);
            if (e) return e;
            if (n1 != normal) return SZO_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH;
            low += s1;
            high += s2;
            if (high > comp) return SZO_ERR_CORRUPTED_JUMPTABLE;
            e = szo_huf_decode_stream(table, cdata + low, (size_t)(high - low), output + normal, (size_t)normal, &n2);
            if (e) return e;
            if (n2 != normal) return SZO_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH;
            low += s2;
            high += s3;
            if (high > regen) return SZO_ERR_PANIC; /* :339-342 panic("Corrupt stream sizes") */
            if (high > comp) return SZO_ERR_CORRUPTED_JUMPTABLE;
            e = szo_huf_decode_stream(table, cdata + low, (size_t)(high - low), output + 2 * normal, (size_t)normal,
                                      &n3);
            if (e) return e;
            if (n3 != normal) return SZO_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH;
            low += s3;
            uint16_t sum16 = (uint16_t)(s1 + s2 + s3);         /* :61 uint16 arithmetic */
            uint16_t s4 = (uint16_t)(comp - (int)sum16);       /* CalcStreamsize4 */
            high += (int)s4;
            if (high != comp) return SZO_ERR_PANIC; /* :356-359 */
            e = szo_huf_decode_stream(table, cdata + low, (size_t)(high - low), output + 3 * normal, (size_t)last, &n4);
            if (e) return e;
            if (n1 + n2 + n3 + n4 != regen) return SZO_ERR_PANIC; /* :366-369 */
        }
    }
    if (ls->type == 2 || ls->type == 3) fs->prev_huf = table; /* carry rule framedecompressor.go:292-294 */
    ls->total_bytes = header_bytes + tree_bytes + comp;      /* framedecompressor.go:103 */
    return SZO_OK;
}

/* one field of SequencesSection.DecodeTables (sequences.go:275-369) */
static int decode_one_table(frame_state *fs, int mode, int field /*0 LL 1 OF 2 ML*/, const uint8_t *src, size_t len,
                            size_t *pos, szo_fse_table **out) {
    szo_fse_table **prev = field == 0 ? &fs->prev_ll : field == 1 ? &fs->prev_of : &fs->prev_ml;
    szo_fse_table *t;
    int e;
    switch (mode) {
    case 0: /* predefined: rebuilt for every block (:281,311,342) */
        t = new_fse(fs);
        if (!t) return SZO_ERR_NOMEM;
        e = field == 0 ? szo_fse_build_ll_table(t) : field == 1 ? szo_fse_build_of_table(t) : szo_fse_build_ml_table(t);
        if (e) return e;
        break;
    case 1: { /* RLE (:282-291, :312-320, :343-351) */
        if (*pos >= len) return SZO_ERR_UNEXPECTED_EOF;
        uint8_t b = src[(*pos)++];
        t = new_fse(fs);
        if (!t) return SZO_ERR_NOMEM;
        t->is_rle = 1;
        if (field == 0) {
            if (b >= 36) return SZO_ERR_PANIC;
            t->rle_value = szo_ll_base[b];
            t->rle_additional_bits = szo_ll_extra[b];
        } else if (field == 1) {
            t->rle_value = b;
            t->rle_additional_bits = 0;
        } else {
            if (b >= 53) return SZO_ERR_PANIC;
            t->rle_value = szo_ml_base[b];
            t->rle_additional_bits = szo_ml_extra[b];
        }
        break;
    }
    case 3: /* repeat (:292-296, :321-325, :352-356) */
        t = *prev;
        if (!t) return field == 0 ? SZO_ERR_NO_LL_TABLE_TO_CARRY_OVER
                                  : field == 1 ? SZO_ERR_NO_OF_TABLE_TO_CARRY_OVER : SZO_ERR_NO_ML_TABLE_TO_CARRY_OVER;
        break;
    default: { /* FSE compressed (:297-305, :326-336, :357-365) */
        int used = 0;
        t = new_fse(fs);
        if (!t) return SZO_ERR_NOMEM;
        e = szo_fse_read_table_description(t, src + *pos, len - *pos, &used);
        if (e) return e;
        *pos += (size_t)used;
        if (field == 0)
            e = szo_fse_build_decoding_table(t, szo_ll_base, 36, szo_ll_extra, 36);
        else if (field == 1)
            e = szo_fse_build_decoding_table(t, NULL, 0, NULL, 0);
        else
            e = szo_fse_build_decoding_table(t, szo_ml_base, 53, szo_ml_extra, 53);
        if (e) return e;
        break;
    }
    }
    *out = t;
    return SZO_OK;
}

/* SequencesSection.DecodeNextSequenceSection (sequences.go:371-450), DecodeSequences
 * (:126-206), DecodeSequence (:64-123).  src/len = the bytes of the block left after the
 * literals section.  *nseq_out sequences are left in fs->seqs; *used_out = header+bitstream
 * bytes consumed. */
static int decode_sequences_section(frame_state *fs, const uint8_t *src, size_t len, int *nseq_out, int *used_out,
                                    int modes_out[3]) {
    size_t pos = 0;
    *nseq_out = 0;
    *used_out = 0;
    modes_out[0] = modes_out[1] = modes_out[2] = -1;
    if (len < 1) return SZO_ERR_UNEXPECTED_EOF;
    uint8_t b0 = src[pos++];
    int need = b0 < 128 ? 1 : (b0 < 255 ? 2 : 3); /* :241-252 */
    if (len < (size_t)need) return SZO_ERR_UNEXPECTED_EOF;
    int nseq; /* :255-269 */
    if (b0 < 128) {
        nseq = b0;
    } else if (b0 < 255) {
        nseq = ((int)(b0 - 128) << 8) + src[pos];
        pos += 1;
    } else {
        nseq = src[pos] + ((int)src[pos + 1] << 8) + 0x7F00;
        pos += 2;
    }
    if (b0 == 0) { /* :395-400 no sequences */
        *used_out = 1;
        return SZO_OK;
    }
    if (pos >= len) return SZO_ERR_UNEXPECTED_EOF;
    uint8_t mb = src[pos++]; /* :406-412 */
    int ll_mode = (mb >> 6) & 3, of_mode = (mb >> 4) & 3, ml_mode = (mb >> 2) & 3;
    modes_out[0] = ll_mode;
    modes_out[1] = of_mode;
    modes_out[2] = ml_mode;
    szo_fse_table *ll, *of, *ml;
    int e;
    e = decode_one_table(fs, ll_mode, 0, src, len, &pos, &ll); /* order LL, OF, ML (:278,308,339) */
    if (e) return e;
    /* the reference assigns each table into the section as it goes; on a later error the
     * already-decoded ones are still carried over by DecodeNextBlockHeader.  Mirror that. */
    fs->prev_ll = ll;
    e = decode_one_table(fs, of_mode, 1, src, len, &pos, &of);
    if (e) return e;
    fs->prev_of = of;
    e = decode_one_table(fs, ml_mode, 2, src, len, &pos, &ml);
    if (e) return e;
    fs->prev_ml = ml;

    const uint8_t *data = src + pos; /* :420-426 the rest is the backward bitstream */
    size_t dlen = len - pos;
    *used_out = (int)len;

    /* DecodeSequences, :126-206 */
    szo_rbits bits;
    szo_rbits_init(&bits, data, dlen);
    int bits_read = 0;
    uint64_t x = 0;
    while (x == 0) { /* :131-139 */
        x = szo_rbits_read(&bits, 1);
        bits_read++;
        if (bits_read > 64) break;
    }
    if (bits_read > 8) return SZO_ERR_BAD_PADDING; /* :141-143 */
    tbl_init_state(ll, &bits);                     /* :145-159 order LL, OF, ML */
    tbl_init_state(of, &bits);
    tbl_init_state(ml, &bits);
    if ((size_t)nseq > fs->seqs_cap) {
        szo_sequence *ns = (szo_sequence *)realloc(fs->seqs, (size_t)nseq * sizeof(*ns));
        if (!ns) return SZO_ERR_NOMEM;
        fs->seqs = ns;
        fs->seqs_cap = (size_t)nseq;
    }
    for (int i = 0; i < nseq; i++) { /* :163-195 */
        int ofcode, llcode, mlcode;
        e = tbl_peek_symbol(of, &ofcode); /* :67-78 peek OF, LL, ML */
        if (e) return e;
        e = tbl_peek_symbol(ll, &llcode);
        if (e) return e;
        e = tbl_peek_symbol(ml, &mlcode);
        if (e) return e;
        if (ofcode > 62 || ofcode < 0) return SZO_ERR_PANIC;
        uint64_t ofx = szo_rbits_read(&bits, ofcode); /* :99-104 */
        fs->seqs[i].offset = (int64_t)(((uint64_t)1 << ofcode) + ofx);
        int mlbits = tbl_additional_bits(ml); /* :106-112 */
        fs->seqs[i].match_length = (int64_t)mlcode + (int64_t)szo_rbits_read(&bits, mlbits);
        int llbits = tbl_additional_bits(ll); /* :114-120 */
        fs->seqs[i].literal_length = (int64_t)llcode + (int64_t)szo_rbits_read(&bits, llbits);
        if (i < nseq - 1) { /* :178-194 update LL, ML, OF; not after the last sequence */
            tbl_next_state(ll, &bits);
            tbl_next_state(ml, &bits);
            tbl_next_state(of, &bits);
        }
    }
    if (szo_rbits_bits_still_in_stream(&bits) != -1) return SZO_ERR_NOT_ALL_BITS_USED; /* :197-204 */
    *nseq_out = nseq;
    return SZO_OK;
}

/* FrameDecompressor.ExecuteSequences, sequence_execution.go:14-63, with
 * LiteralSection.Read / GetRest (literals.go:383-420) inlined. */
static int execute_sequences(frame_state *fs, lit_section *ls, int nseq, outbuf *o, szo_trace *tr) {
    for (int i = 0; i < nseq; i++) {
        const szo_sequence *seq = &fs->seqs[i];
        if (seq->literal_length > 0) { /* :19-34 */
            int64_t ll = seq->literal_length;
            if (ll > MAX_BLOCK) return SZO_ERR_PANIC; /* literalsCopyBuf[:ll] */
            int e = out_reserve(o, (size_t)ll);
            if (e) return e;
            if (ls->type == 1) { /* literals.go:390-396 RLE literals never run dry */
                if (ls->data_len < 1) return SZO_ERR_PANIC;
                memset(o->buf + o->len, ls->data[0], (size_t)ll);
                ls->data_read += (int)ll;
            } else {
                if (ls->data_read == ls->data_len) return SZO_ERR_UNEXPECTED_EOF; /* literals.go:398-400 io.EOF */
                int64_t avail = ls->data_len - ls->data_read;
                if (avail < ll) return SZO_ERR_DIDNT_COPY_ALL_LITERAL_BYTES; /* :26-28 */
                memcpy(o->buf + o->len, ls->data + ls->data_read, (size_t)ll);
                ls->data_read += (int)ll;
            }
            o->len += (size_t)ll;
        }
        int64_t offset = szo_next_offset(fs->offset_history, seq); /* :43 */
        if (tr) tr->real_offsets[tr->nsequences - (size_t)nseq + (size_t)i] = offset;
        if (seq->match_length > 0) { /* :44-49 */
            int e = out_reserve(o, (size_t)seq->match_length);
            if (e) return e;
            e = szo_match_copy(o->buf, &o->len, o->cap, seq->match_length, offset);
            if (e) return e;
        }
    }
    /* :55-60 trailing literals = GetRest (literals.go:411-420) */
    if (ls->type == 1) {
        int64_t rest = (int64_t)ls->regen - ls->data_read;
        if (rest < 0) return SZO_ERR_PANIC; /* make([]byte, negative) */
        if (rest > 0 && ls->data_len < 1) return SZO_ERR_PANIC;
        int e = out_reserve(o, (size_t)rest);
        if (e) return e;
        if (rest > 0) memset(o->buf + o->len, ls->data[0], (size_t)rest);
        o->len += (size_t)rest;
    } else {
        int64_t rest = ls->data_len - ls->data_read;
        int e = out_reserve(o, (size_t)rest);
        if (e) return e;
        memcpy(o->buf + o->len, ls->data + ls->data_read, (size_t)rest);
        o->len += (size_t)rest;
    }
    return SZO_OK;
}

/* FrameDecompressor.Decompress (framedecompressor.go:153-170): CheckMagicnum :130-150,
 * DecodeFrameHeader :306-374, decodeAllBlocks :246-267, DecodeNextBlock :198-244,
 * DecodeNextBlockHeader :270-303 + Block.DecodeHeader (block.go:33-55),
 * DecodeNextBlockContent :93-126. */
/* ------------------------------------------------------------------------- */
/* Dictionaries.  NOT in the reference (Readme.md:59-61 lists them as missing; frame.go:38-47 parses the
 * Dictionary_ID and ignores it): this part restates the zstd format specification (RFC 8878 section 5, "Dictionary Format"):
 *   raw-content dictionary:  any bytes; they are the history in front of the frame (matches may reach into them);
 *   formatted dictionary:    magic 0xEC30A437 | Dictionary_ID (4, LE) | Huffman tree description for literals |
 *                            FSE table descriptions for offsets, match lengths, literal lengths (that order) |
 *                            three repeat offsets (4 bytes LE each) | content.
 * The tables act as the "previous block" of the frame's first block (Treeless literals, Repeat modes), the repeat offsets
 * replace 1, 4, 8.  Pinned by frames libzstd 1.5.5 compressed WITH the dictionary (tools/corpusgen.c: ZSTD_compress_usingDict,
 * ZDICT_trainFromBuffer): the oracle must give back the original bytes (tests/test_dictionary.py). */
static int load_dictionary(frame_state *fs, const uint8_t *dict, size_t dlen, const uint8_t **content, size_t *content_len,
                           uint32_t *dict_id) {
    *content = dict;
    *content_len = dlen;
    *dict_id = 0;
    if (dlen < 8 || !(dict[0] == 0x37 && dict[1] == 0xA4 && dict[2] == 0x30 && dict[3] == 0xEC)) return SZO_OK; /* raw content */
    *dict_id = (uint32_t)dict[4] | ((uint32_t)dict[5] << 8) | ((uint32_t)dict[6] << 16) | ((uint32_t)dict[7] << 24);
    size_t pos = 8;
    {   /* Huffman tree description, as in a Compressed literals section (literals.go:254-267) */
        uint8_t weights[4096];
        int nweights = 0, tree_bytes = 0;
        int e = szo_huf_decode_tree_desc(dict + pos, dlen - pos, weights, &nweights, &tree_bytes);
        if (e) return e;
        pos += (size_t)tree_bytes;
        szo_huf_table *table = (szo_huf_table *)calloc(1, sizeof(*table));
        if (!table) return SZO_ERR_NOMEM;
        if (own(fs, table, 1)) {
            free(table);
            return SZO_ERR_NOMEM;
        }
        e = szo_huf_build(weights, nweights, table);
        if (e) return e;
        fs->prev_huf = table;
    }
    for (int k = 0; k < 3; k++) { /* offsets, match lengths, literal lengths */
        int used = 0, e;
        szo_fse_table *t = new_fse(fs);
        if (!t) return SZO_ERR_NOMEM;
        e = szo_fse_read_table_description(t, dict + pos, dlen - pos, &used);
        if (e) return e;
        pos += (size_t)used;
        if (k == 0) {
            e = szo_fse_build_decoding_table(t, NULL, 0, NULL, 0);
            fs->prev_of = t;
        } else if (k == 1) {
            e = szo_fse_build_decoding_table(t, szo_ml_base, 53, szo_ml_extra, 53);
            fs->prev_ml = t;
        } else {
            e = szo_fse_build_decoding_table(t, szo_ll_base, 36, szo_ll_extra, 36);
            fs->prev_ll = t;
        }
        if (e) return e;
    }
    if (dlen - pos < 12) return SZO_ERR_UNEXPECTED_EOF;
    for (int k = 0; k < 3; k++) {
        uint32_t r = (uint32_t)dict[pos] | ((uint32_t)dict[pos + 1] << 8) | ((uint32_t)dict[pos + 2] << 16) | ((uint32_t)dict[pos + 3] << 24);
        pos += 4;
        fs->offset_history[k] = r;
    }
    for (int k = 0; k < 3; k++) /* a repeat offset of the dictionary points into its content */
        if (fs->offset_history[k] == 0 || (uint64_t)fs->offset_history[k] > dlen - pos) return SZO_ERR_CANT_REPEAT_BYTES;
    *content = dict + pos;
    *content_len = dlen - pos;
    return SZO_OK;
}

static int decode_frame_impl(const uint8_t *src, size_t len, const uint8_t *dict, size_t dict_len, uint8_t **out, size_t *out_len,
                             szo_trace *tr);

int szo_decode_frame(const uint8_t *src, size_t len, uint8_t **out, size_t *out_len, szo_trace *tr) {
    return decode_frame_impl(src, len, NULL, 0, out, out_len, tr);
}
int szo_decode_frame_dict(const uint8_t *src, size_t len, const uint8_t *dict, size_t dict_len, uint8_t **out, size_t *out_len,
                          szo_trace *tr) {
    return decode_frame_impl(src, len, dict, dict_len, out, out_len, tr);
}

static int decode_frame_impl(const uint8_t *src, size_t len, const uint8_t *dict, size_t dict_len, uint8_t **out, size_t *out_len,
                             szo_trace *tr) {
    frame_state fs;
    outbuf o = {NULL, 0, 0};
    size_t pos = 0;
    size_t prefix = 0; /* dictionary content in front of the frame's output */
    int rc = SZO_OK;
    memset(&fs, 0, sizeof(fs));
    if (tr) memset(tr, 0, sizeof(*tr));
    *out = NULL;
    *out_len = 0;
    fs.offset_history[0] = 1; /* framedecompressor.go:48,59 */
    fs.offset_history[1] = 4;
    fs.offset_history[2] = 8;
    if (dict && dict_len) {
        const uint8_t *content;
        size_t clen;
        uint32_t id;
        rc = load_dictionary(&fs, dict, dict_len, &content, &clen, &id);
        if (rc) {
            frame_state_free(&fs);
            return rc;
        }
        if (clen) {
            if (out_reserve(&o, clen)) {
                frame_state_free(&fs);
                return SZO_ERR_NOMEM;
            }
            memcpy(o.buf, content, clen);
            o.len = prefix = clen;
        }
    }

    /* CheckMagicnum */
    if (len < 4) {
        rc = SZO_ERR_UNEXPECTED_EOF;
        goto done;
    }
    if (!(src[0] == 0x28 && src[1] == 0xB5 && src[2] == 0x2F && src[3] == 0xFD)) {
        rc = SZO_ERR_WRONG_MAGICNUMBER;
        goto done;
    }
    pos = 4;

    /* DecodeFrameHeader + frame.go getters */
    if (pos >= len) {
        rc = SZO_ERR_UNEXPECTED_EOF;
        goto done;
    }
    uint8_t fhd = src[pos++];
    int single_segment = (fhd >> 5) & 1;                   /* frame.go:101-103 */
    int dict_flag = fhd & 3;                               /* frame.go:113-127 */
    int dict_size = dict_flag == 3 ? 4 : dict_flag;
    int fcs_flag = fhd >> 6;                               /* frame.go:79-98 */
    int fcs_size = fcs_flag == 0 ? (single_segment ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    int header_size = (single_segment ? 0 : 1) + dict_size + fcs_size;
    if (len - pos < (size_t)header_size) {
        rc = SZO_ERR_UNEXPECTED_EOF;
        goto done;
    }
    uint64_t window_size = 0, fcs = 0;
    if (!single_segment) { /* frame.go:28-36 */
        uint8_t wd = src[pos++];
        unsigned exp = wd >> 3;
        uint64_t mant = wd & 7;
        uint64_t base = (uint64_t)1 << (10 + exp);
        window_size = base + (base / 8) * mant;
    }
    pos += (size_t)dict_size; /* dictionary id is parsed and ignored (frame.go:38-47) */
    if (fcs_size > 0) {      /* frame.go:49-61 */
        for (int i = 0; i < fcs_size; i++) fcs |= (uint64_t)src[pos + (size_t)i] << (8 * i);
        if (fcs_size == 2) fcs += 256;
        pos += (size_t)fcs_size;
        if (single_segment) window_size = fcs; /* framedecompressor.go:358-360 */
    }
    if (tr) {
        tr->window_size = window_size;
        tr->frame_content_size = fcs;
        tr->has_fcs = fcs_size > 0;
        tr->single_segment = single_segment;
    }
    if (fcs_size > 0 && fcs < ((uint64_t)1 << 40)) {
        if (out_reserve(&o, (size_t)fcs + 1)) { /* room beyond what is there already (the dictionary's content) */
            rc = SZO_ERR_NOMEM;
            goto done;
        }
    }
    fs.lit_data = (uint8_t *)malloc(MAX_BLOCK);
    if (!fs.lit_data) {
        rc = SZO_ERR_NOMEM;
        goto done;
    }

    /* decodeAllBlocks */
    int last_block = 0;
    while (!last_block) {
        /* DecodeNextBlockHeader */
        if (len - pos < 3) {
            rc = SZO_ERR_UNEXPECTED_EOF;
            goto done;
        }
        const uint8_t *h = src + pos;
        pos += 3;
        last_block = h[0] & 1;         /* block.go:38 */
        int btype = (h[0] >> 1) & 3;   /* block.go:39 */
        uint64_t bsize = (uint64_t)(h[0] >> 3) + ((uint64_t)h[1] << 5) + ((uint64_t)h[2] << 13);
        if (btype >= 3) {
            rc = SZO_ERR_ILLEGAL_BLOCK_TYPE;
            goto done;
        }
        if (bsize > MAX_BLOCK) {
            rc = SZO_ERR_ILLEGAL_BLOCK_SIZE;
            goto done;
        }
        szo_block_trace *bt = NULL;
        if (tr) {
            bt = trace_new_block(tr);
            if (!bt) {
                rc = SZO_ERR_NOMEM;
                goto done;
            }
            bt->type = btype;
            bt->last = last_block;
            bt->block_size = (uint32_t)bsize;
            bt->out_off = o.len - prefix;
        }
        if (btype == 0) { /* Raw: framedecompressor.go:211-215 */
            if (len - pos < bsize) {
                rc = SZO_ERR_UNEXPECTED_EOF;
                goto done;
            }
            if ((rc = out_reserve(&o, (size_t)bsize))) goto done;
            memcpy(o.buf + o.len, src + pos, (size_t)bsize);
            o.len += (size_t)bsize;
            pos += (size_t)bsize;
        } else if (btype == 1) { /* RLE: framedecompressor.go:229-241 (reads its byte even when size is 0) */
            if (len - pos < 1) {
                rc = SZO_ERR_UNEXPECTED_EOF;
                goto done;
            }
            uint8_t b = src[pos++];
            if ((rc = out_reserve(&o, (size_t)bsize))) goto done;
            memset(o.buf + o.len, b, (size_t)bsize);
            o.len += (size_t)bsize;
        } else { /* Compressed: DecodeNextBlockContent + ExecuteSequences */
            if (len - pos < bsize) {
                rc = SZO_ERR_UNEXPECTED_EOF; /* LimitedReader over a short source */
                goto done;
            }
            const uint8_t *bp = src + pos;
            lit_section ls;
            rc = decode_literals_section(&fs, bp, (size_t)bsize, &ls);
            if (rc) goto done;
            int64_t bytes_left = (int64_t)bsize - ls.total_bytes; /* framedecompressor.go:103-104 */
            if (bytes_left < 0) {
                rc = SZO_ERR_PANIC;
                goto done;
            }
            int nseq = 0, used = 0, modes[3];
            rc = decode_sequences_section(&fs, bp + ls.total_bytes, (size_t)bytes_left, &nseq, &used, modes);
            if (rc) goto done;
            if ((int64_t)ls.total_bytes + used != (int64_t)bsize) { /* :114-123 */
                rc = SZO_ERR_CORRUPT_SIZES;
                goto done;
            }
            if (tr) {
                bt->lit_type = ls.type;
                bt->lit_streams = ls.streams;
                bt->lit_regen = (uint32_t)ls.regen;
                bt->lit_off = tr->nliterals;
                bt->huf_max_bits = ls.huf_max_bits;
                if (ls.regen > 0) {
                    if (ls.type == 1 && ls.data_len < 1) {
                        rc = SZO_ERR_PANIC;
                        goto done;
                    }
                    if ((rc = trace_add_literals(tr, ls.data, (size_t)ls.regen, ls.type == 1))) goto done;
                }
                bt->nseq = (uint32_t)nseq;
                bt->seq_off = tr->nsequences;
                bt->ll_mode = modes[0];
                bt->of_mode = modes[1];
                bt->ml_mode = modes[2];
                if ((rc = trace_reserve_seq(tr, (size_t)nseq))) goto done;
                if (nseq) memcpy(tr->sequences + tr->nsequences, fs.seqs, (size_t)nseq * sizeof(szo_sequence));
                tr->nsequences += (size_t)nseq;
            }
            rc = execute_sequences(&fs, &ls, nseq, &o, tr);
            if (rc) goto done;
            pos += (size_t)bsize;
        }
        if (bt) {
            bt->out_len = o.len - prefix - bt->out_off;
            bt->hist_after[0] = fs.offset_history[0];
            bt->hist_after[1] = fs.offset_history[1];
            bt->hist_after[2] = fs.offset_history[2];
        }
    }
    if (tr) tr->bytes_consumed = pos; /* the optional 4-byte checksum is never read */
done:
    frame_state_free(&fs);
    if (rc != SZO_OK) {
        free(o.buf);
        return rc;
    }
    if (!o.buf) o.buf = (uint8_t *)malloc(1);
    if (prefix) memmove(o.buf, o.buf + prefix, o.len - prefix); /* the frame's own bytes */
    *out = o.buf;
    *out_len = o.len - prefix;
    return SZO_OK;
}

/* ------------------------------------------------------------------------- */
/* multi-threaded batch driver for the CPU baseline                           */
typedef struct {
    const uint8_t *src;
    const uint64_t *frame_off, *frame_len, *dst_off;
    uint32_t nframes;
    uint8_t *dst;
    size_t dst_cap;
    uint64_t *out_len;
    int32_t *status;
    volatile uint32_t *next;
} mt_job;

static void *mt_worker(void *arg) {
    mt_job *j = (mt_job *)arg;
    for (;;) {
        uint32_t i = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->nframes) break;
        uint8_t *o = NULL;
        size_t n = 0;
        int rc = szo_decode_frame(j->src + j->frame_off[i], (size_t)j->frame_len[i], &o, &n, NULL);
        if (rc == SZO_OK && j->dst && j->dst_off) {
            if (j->dst_off[i] + n <= j->dst_cap)
                memcpy(j->dst + j->dst_off[i], o, n);
            else
                rc = SZO_ERR_NOMEM;
        }
        if (j->out_len) j->out_len[i] = n;
        if (j->status) j->status[i] = rc;
        free(o);
    }
    return NULL;
}

int szo_decode_batch_mt(const uint8_t *src, const uint64_t *frame_off, const uint64_t *frame_len, uint32_t nframes,
                        uint8_t *dst, size_t dst_cap, const uint64_t *dst_off, uint64_t *out_len, int32_t *status,
                        int nthreads) {
    volatile uint32_t next = 0;
    mt_job job = {src, frame_off, frame_len, dst_off, nframes, dst, dst_cap, out_len, status, &next};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    if (!th) return SZO_ERR_NOMEM;
    int started = 0;
    for (int t = 0; t < nthreads; t++) {
        if (pthread_create(&th[t], NULL, mt_worker, &job) != 0) break;
        started++;
    }
    if (started == 0) mt_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    return SZO_OK;
}
