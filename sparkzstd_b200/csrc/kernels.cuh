// kernels.cuh -- the four GPU stages of the decode path (sm_100a).
//
//   k_build_huf_tables  stage 1: Huffman tree description (direct / FSE-compressed weights) and
//                       decode table in shared memory, one warp per tree, published to HBM
//                       (structure/huffman.go:40-190)
//   k_decode_literals   stage 2: 1- or 4-stream literal decode, one LANE per stream, 8 blocks per warp
//                       (structure/huffman.go:192-264, structure/literals.go:209-373)
//   k_build_seq_tables  stage 1: LL/OF/ML table selection + construction in shared memory, one
//                       warp per block, published to a table arena in HBM
//                       (fse/fse.go, fse/predefined.go, structure/sequences.go:275-369)
//   k_decode_sequences  stage 3: three interleaved FSE states over the backward bitstream, one
//                       LANE per block, tables resident in shared memory
//                       (structure/sequences.go:64-206)
//   k_scan_blocks       device-wide exclusive prefix sum of per-block regenerated sizes
//                       (the reference gets positions for free from its ring buffer,
//                       decompression/ringbuffer.go:102-178)
//   k_frame_verdict     per frame: first failing block, placement in the output
//   k_execute_bodies    Raw / RLE block bodies and blocks without sequences, one warp each
//                       (framedecompressor.go:211-241)
//   k_execute           stage 4, one warp per frame: repeat-offset history and positions by warp scans,
//                       then the output produced in address order, 128 bytes per step, every lane
//                       fetching the source byte of its output byte (literal or match)
//                       (decompression/sequence_execution.go:14-114, ringbuffer.go:197-277)
//   k_execute_pair      the same with a producer warp and a consumer warp per frame: the long frames the
//                       block-parallel path cannot take
//   k_long_hist / k_long_compose / k_long_emit / k_long_jump / k_long_verdict  (execute_long.cuh)
//                       stage 4 of LONG frames, parallel over their blocks and bytes: repeat-offset history by
//                       composing per-block transfer functions, then one distance cell per output byte resolved
//                       by in-place pointer jumping
//   k_verify_checksums  optional XXH64 content checksum, one thread per frame
//
// All arithmetic is integer; there is no tensor-core work on this path.
#pragma once
#include <cuda_runtime.h>

#include "batch.cuh"
#include "huffman.cuh"
#include "sequences.cuh"

namespace szb {

#ifndef SZB_SERIAL_TABLES
#define SZB_SERIAL_TABLES 0
#endif


// ---------------------------------------------------------------------------------------------
// Builds one FSE decode table into `table` from `ts`.  Lane 0 parses, the warp builds.
// Returns (warp-uniform) status; *al_out = accuracy log (0 for RLE), *used_out = table bytes.
__device__ __forceinline__ int build_fse_table(const TableSource &ts, int kind, const uint32_t *predef, uint32_t *table,
                                               int16_t *norm, uint16_t *next, uint8_t *symk, uint32_t *al_out,
                                               uint32_t *used_out) {
    const uint32_t lane = threadIdx.x & 31;
    if (ts.mode == 0) {  // predefined (sequences.go:279-281,309-311,340-342); the cells were built once on the host
        const uint32_t off = kind == KIND_LL ? 0 : (kind == KIND_OF ? 64 : 96);
        const uint32_t n = kind == KIND_OF ? 32 : 64;
        for (uint32_t i = lane; i < n; i += 32) table[i] = predef[off + i];
        *al_out = kind == KIND_OF ? 5 : 6;
        *used_out = 0;
        __syncwarp();
        return SZB_OK;
    }
    if (ts.mode == 1) {  // RLE (sequences.go:282-291,312-320,343-351): a one-cell table, every step reads 0 bits
        int rc = SZB_OK;
        if (lane == 0) {
            if (ts.avail < 1) {
                rc = SZB_ERR_UNEXPECTED_EOF;
            } else {
                uint32_t code = ts.p[0];
                if ((kind == KIND_LL && code >= 36) || (kind == KIND_ML && code >= 53))
                    rc = SZB_ERR_PANIC;  // index out of range in the reference
                else if (kind == KIND_OF && code > 31)
                    rc = SZB_ERR_UNSUPPORTED;
                else
                    table[0] = fse_pack(0, 0, extra_bits_for(kind, code), code);
            }
        }
        rc = __shfl_sync(kFull, rc, 0);
        *al_out = 0;
        *used_out = 1;
        __syncwarp();
        return rc;
    }
    // FSE-compressed (sequences.go:297-305,326-336,357-365)
    uint32_t nsym = 0, al = 0, used = 0;
    int rc = SZB_OK;
    if (lane == 0) {
        uint32_t max_al = kind == KIND_OF ? kMaxALOF : (kind == KIND_LL ? kMaxALLL : (kind == KIND_ML ? kMaxALML : kMaxALHufW));
        rc = fse_read_description(ts.p, ts.avail, max_al, norm, &nsym, &al, &used);
        if (rc == SZB_OK && kind == KIND_OF && nsym > 32) rc = SZB_ERR_UNSUPPORTED;
    }
    rc = __shfl_sync(kFull, rc, 0);
    if (rc) return rc;
    nsym = __shfl_sync(kFull, nsym, 0);
    al = __shfl_sync(kFull, al, 0);
    used = __shfl_sync(kFull, used, 0);
    __syncwarp();
#if SZB_SERIAL_TABLES
    if (lane == 0) rc = fse_build_serial(norm, nsym, al, kind, table, next);
    rc = __shfl_sync(kFull, rc, 0);
#else
    rc = fse_build_warp(norm, nsym, al, kind, table, next, symk);
#endif
    __syncwarp();
    *al_out = al;
    *used_out = used;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Huffman decode table, warp-cooperative fill (same cells as huf_build_serial).
__device__ __forceinline__ int huf_build_warp(const uint8_t *weights, uint32_t nw, uint16_t *table, uint8_t *sorted,
                                              uint32_t *max_bits_out) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t rank_count[kMaxHufBits + 2];
    uint32_t max_bits = 0, last_nb = 0;
    int rc = SZB_OK;
    // weight statistics are cheap and uniform: every lane computes them redundantly
    rc = huf_weight_stats(weights, nw, &max_bits, &last_nb, rank_count);
    if (rc) return rc;
    const uint32_t size = 1u << max_bits;
    // rank_cell[nb] = first cell of bit length nb; rank_sym[nb] = first slot in `sorted`
    uint32_t rank_cell[kMaxHufBits + 2], rank_sym[kMaxHufBits + 2];
    {
        uint32_t cell = 0, sym = 0;
        for (uint32_t nb = max_bits; nb >= 1; nb--) {
            rank_cell[nb] = cell;
            rank_sym[nb] = sym;
            cell += rank_count[nb] << (max_bits - nb);
            sym += rank_count[nb];
        }
        if (cell != size) return cell > size ? SZB_ERR_PANIC : SZB_ERR_CORRUPTED_HUFF_TREE;  // huffman.go:173-175
        rank_cell[0] = cell;
    }
    // counting sort of the symbols by (bit length, symbol): 32 symbols per step
    uint32_t run[kMaxHufBits + 2];
    for (uint32_t nb = 0; nb <= max_bits; nb++) run[nb] = 0;
    for (uint32_t s0 = 0; s0 <= nw; s0 += 32) {
        uint32_t s = s0 + lane;
        uint32_t nb = 0;
        if (s < nw) {
            uint32_t w = weights[s];
            nb = w ? max_bits + 1 - w : 0;
        } else if (s == nw) {
            nb = last_nb;
        }
        for (uint32_t r = 1; r <= max_bits; r++) {  // uniform loop, <= 11 ballots
            uint32_t m = __ballot_sync(kFull, nb == r);
            if (nb == r) sorted[rank_sym[r] + run[r] + __popc(m & ((1u << lane) - 1))] = (uint8_t)s;
            run[r] += __popc(m);
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < size; i += 32) {
        uint32_t nb = max_bits;
        while (nb > 1 && i >= rank_cell[nb - 1]) nb--;  // cells are grouped by descending bit length
        uint32_t k = (i - rank_cell[nb]) >> (max_bits - nb);
        table[i] = (uint16_t)(sorted[rank_sym[nb] + k] | (nb << 8));
    }
    __syncwarp();
    *max_bits_out = max_bits;
    return SZB_OK;
}

// shared memory per warp, k_build_huf_tables
struct HufSmem {
    union {
        uint16_t huf[1 << kMaxHufBits];  // 4 KB
        uint32_t fse[1 << kMaxALHufW];   // 2 KB, dead once the weights are decoded
    } t;
    uint8_t weights[256];
    uint8_t sorted[256];
    uint8_t symk[1 << kMaxALHufW];
    int16_t norm[kMaxFseSymbols];
    uint16_t next[kMaxFseSymbols];
};

// Stage 1 for literals: one warp per block that CARRIES a Huffman tree description (literals type
// Compressed) decodes the weights (direct or FSE-compressed, huffman.go:40-107), builds the decode
// table in shared memory (huffman.go:112-190) and publishes it to the table arena in HBM.  Treeless
// blocks (literals.go:247-252) later read the table of their origin block: it is built once, not
// once per user.
__global__ void __launch_bounds__(kCtaThreads) k_build_huf_tables(DeviceBatch a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kWarpsPerCta + warp_in_cta;
    if (w >= a.n_hufo) return;
    HufSmem &sm = reinterpret_cast<HufSmem *>(smem_raw)[warp_in_cta];
    const szb_block_desc o = a.blocks[a.hufo_list[w]];

    // --- tree description (huffman.go:40-107) ---
    const uint8_t *tree = a.src + o.src_off + o.lit_hdr_bytes;
    const uint32_t tree_avail = o.lit_comp;
    int rc = SZB_OK;
    uint32_t nw = 0, tree_bytes = 0;
    uint32_t hb = 0;
    if (tree_avail < 1)
        rc = SZB_ERR_UNEXPECTED_EOF;
    else
        hb = tree[0];
    if (rc == SZB_OK) {
        if (hb < 128) {  // FSE-compressed weights
            TableSource ts{tree + 1, tree_avail - 1, 2};
            uint32_t al = 0, used = 0;
            if (1 + hb > tree_avail) rc = SZB_ERR_UNEXPECTED_EOF;
            if (rc == SZB_OK) rc = build_fse_table(ts, KIND_HUFW, a.predef, sm.t.fse, sm.norm, sm.next, sm.symk, &al, &used);
            if (rc == SZB_OK) {
                if (lane == 0) {
                    if (used > hb)
                        rc = SZB_ERR_PANIC;  // make([]byte, negative), huffman.go:67-68
                    else
                        rc = fse_decode_weights(sm.t.fse, al, tree + 1 + used, hb - used, sm.weights, &nw);
                }
                rc = __shfl_sync(kFull, rc, 0);
                nw = __shfl_sync(kFull, nw, 0);
            }
            tree_bytes = 1 + hb;
        } else {  // direct weights
            nw = hb - 127;
            if (lane == 0) rc = huf_read_direct_weights(tree + 1, tree_avail - 1, nw, sm.weights);
            rc = __shfl_sync(kFull, rc, 0);
            tree_bytes = 1 + ((nw + 1) >> 1);
        }
    }
    __syncwarp();
    uint32_t max_bits = 0;
    if (rc == SZB_OK) {
#if SZB_SERIAL_TABLES
        if (lane == 0) rc = huf_build_serial(sm.weights, nw, sm.t.huf, &max_bits);
        rc = __shfl_sync(kFull, rc, 0);
        max_bits = __shfl_sync(kFull, max_bits, 0);
#else
        rc = huf_build_warp(sm.weights, nw, sm.t.huf, sm.sorted, &max_bits);
#endif
    }
    __syncwarp();
    if (rc == SZB_OK) {
        uint16_t *slot = a.huf_tabs + (size_t)w * (1u << kMaxHufBits);
        for (uint32_t i = lane; i < (1u << max_bits); i += 32) slot[i] = sm.t.huf[i];
    }
    if (lane == 0) {
        HufInfo info;
        info.max_bits = (uint8_t)max_bits;
        info.pad = 0;
        info.tree_bytes = (uint16_t)tree_bytes;
        info.status = rc;
        a.huf_info[w] = info;
    }
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g));
}

// Reverse bit reader for one lane of k_decode_literals: the 64-bit window of bits.cuh, refilled with
// 32-bit words from a per-lane shared-memory ring of four 16-byte chunks of the stream.  The ring is
// topped up with cp.async once per group of eight symbols, by all lanes at the same point of the
// program (the lanes walk unrelated streams in lock step), one group ahead of use, so that no lane waits
// for HBM while its neighbours are ready.
//
// Eight symbols are at most 88 bits: the refills of one group move at most four words (16 bytes) out of
// the ring.  The top-up of group g requests every chunk down to the one holding byte (next - kHufRingAhead)
// and is waited for at group g+1 (wait_group 1), whose reads stay above next_g - 32: kHufRingAhead = 32.
// The ring then spans chunk(next - 32) .. chunk(next - 1), three of its four slots, so a copy never
// lands in a slot that is still being read.
constexpr uint32_t kHufRingBytes = 64;
constexpr uint32_t kHufRingStride = 80;  // 64 B ring + pad (keeps 16-byte alignment, spreads banks)
constexpr int32_t kHufRingAhead = 32;
struct HufBits {
    uint64_t win;
    int32_t avail;
    int32_t next;        // stream bytes [0, next) not yet moved into the window
    int32_t remaining;   // real bits not consumed yet
    int32_t lowreq;      // lowest chunk (16 B units from the aligned address at or below the stream) requested so far
    uint32_t mis;        // stream address & 15
    uint32_t ring_saddr; // shared-space address of this lane's ring
    const uint8_t *ring;
    const uint8_t *base;
    const uint4 *chunk0;
};

__device__ __forceinline__ void hring_fetch_if(const HufBits &r, bool p, int32_t c) {
    const uint32_t dst = r.ring_saddr + (((uint32_t)c & (kHufRingBytes / 16 - 1)) << 4);
    const uint4 *src = r.chunk0 + c;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %0, 0;\n\t"
        "@p cp.async.cg.shared.global [%1], [%2], 16;\n\t}"
        ::"r"((uint32_t)p), "r"(dst), "l"(src)
        : "memory");
}
// steady state, no branch: at most two more chunks, which is what a group can consume
__device__ __forceinline__ void hring_topup(HufBits &r) {
    int32_t need = ((int32_t)r.mis + r.next - kHufRingAhead) >> 4;
    need = need < 0 ? 0 : need;
    const bool p1 = r.lowreq - 1 >= need, p2 = r.lowreq - 2 >= need;
    hring_fetch_if(r, p1, r.lowreq - 1);
    hring_fetch_if(r, p2, r.lowreq - 2);
    r.lowreq -= (int32_t)p1 + (int32_t)p2;
    asm volatile("cp.async.commit_group;");
}
// everything down to (next - kHufRingAhead), and wait for it: stream start, and around the unaligned head and tail
__device__ __forceinline__ void hring_fill(HufBits &r) {
    int32_t need = ((int32_t)r.mis + r.next - kHufRingAhead) >> 4;
    need = need < 0 ? 0 : need;
    while (r.lowreq > need) hring_fetch_if(r, true, --r.lowreq);
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void hring_refill(HufBits &r) {
    if (r.avail <= 32) {
        if (r.next >= 4) {
            const uint32_t wa = r.mis + (uint32_t)r.next - 4;  // word position relative to chunk 0
            const uint32_t w = *reinterpret_cast<const uint32_t *>(r.ring + (wa & (kHufRingBytes - 1)));
            r.next -= 4;
            r.win |= (uint64_t)w << (32 - r.avail);
            r.avail += 32;
        } else {
            while (r.next > 0) {
                const uint32_t b = r.base[--r.next];
                r.win |= (uint64_t)b << (56 - r.avail);
                r.avail += 8;
            }
            r.avail = 64;  // zero fill below the stream start (reversebitstream.go:23-27,67-75)
        }
    }
}

__device__ __forceinline__ bool hring_init(HufBits &r, const uint8_t *data, int32_t len, uint8_t *ring) {
    r.base = data;
    r.next = len;
    r.avail = 0;
    r.win = 0;
    r.remaining = len * 8;
    r.ring = ring;
    r.ring_saddr = (uint32_t)__cvta_generic_to_shared(ring);
    r.mis = (uint32_t)(reinterpret_cast<uintptr_t>(data) & 15);
    r.chunk0 = reinterpret_cast<const uint4 *>(data - r.mis);
    r.lowreq = 0;
    if (len <= 0) {
        r.avail = 64;
        return false;
    }
    r.lowreq = (int32_t)((r.mis + (uint32_t)len - 1) >> 4) + 1;
    // byte loads until the unread length is a multiple of 4 from an aligned address
    while (r.next > 0 && ((r.mis + (uint32_t)r.next) & 3) != 0) {
        const uint32_t b = r.base[--r.next];
        r.win |= (uint64_t)b << (56 - r.avail);
        r.avail += 8;
    }
    if (r.next == 0) r.avail = 64;
    hring_fill(r);
    hring_refill(r);
    return true;
}

__device__ __forceinline__ uint32_t hring_peek(const HufBits &r, uint32_t n) { return (uint32_t)((r.win >> 1) >> (63 - n)); }
__device__ __forceinline__ void hring_skip(HufBits &r, uint32_t n) {
    r.win <<= n;
    r.avail -= (int32_t)n;
    r.remaining -= (int32_t)n;
}

// 16 symbols of one stream into one aligned 16-byte store (the body of huf_decode_stream_vec's main loop).
template <bool kTrack>
__device__ __forceinline__ void huf_group16(HufBits &r, uint32_t tb, uint32_t psh, uint8_t *dst, bool &clean) {
    uint32_t wv[4];
    uint32_t used = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if ((q & 1) == 0) {  // a group of eight symbols
            hring_topup(r);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        }
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            if ((t & 1) == 0) hring_refill(r);
            uint32_t e;
            asm("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(tb + 2 * ((uint32_t)(r.win >> 32) >> psh)));
            acc |= (e & 0xFF) << (8 * t);
            const uint32_t nb = e >> 8;
            r.win <<= nb;
            r.avail -= (int32_t)nb;
            if (kTrack) {
                r.remaining -= (int32_t)nb;
                clean |= r.remaining == 0;
            } else {
                used += nb;
            }
        }
        wv[q] = acc;
    }
    if (!kTrack) r.remaining -= (int32_t)used;
    *reinterpret_cast<uint4 *>(dst) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
}

// HuffmanDecodingTable.DecodeStream (huffman.go:221-264) for one stream, one lane; same recurrence as
// huf_decode_stream (huffman.cuh) with the symbols buffered 16 deep in registers so they leave as
// aligned 16-byte stores, and one refill check per two symbols (2 x 11 bits <= the 32 guaranteed).
__device__ __forceinline__ int huf_decode_stream_vec(const uint16_t *table, uint32_t max_bits, const uint8_t *p, uint32_t len,
                                                     uint8_t *out, uint32_t expected, uint8_t *ring, bool one_of_four) {
    HufBits r;
    if (!hring_init(r, p, (int32_t)len, ring)) return SZB_ERR_BAD_PADDING;
    {   // skip padding: zero bits then the first 1 bit, at most 8 (huffman.go:227-238)
        uint32_t top = (uint32_t)(r.win >> 56);
        if (top == 0) return SZB_ERR_BAD_PADDING;
        hring_skip(r, (uint32_t)__clz(top) - 24 + 1);
    }
    // The reference decodes symbols while bits remain and then looks at how the stream ended (huffman.go:248-261): exactly
    // used up -> fine (too few symbols is the caller's ErrStreamDidntDecodeToRightLength), overrun -> ErrDidntUseAllBits.  Here
    // `expected` symbols are decoded whatever the stream holds; `clean` remembers whether the bits ever ran out exactly at a
    // symbol boundary, which is where the reference would have stopped.
    bool clean = r.remaining == 0;
    uint32_t n = 0;
    uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15)) & 15;
    if (head > expected) head = expected;
    // up to 15 single symbols (165 bits) to reach a 16-byte aligned output address: within what hring_init made resident
    for (; n < head; n++) {
        hring_refill(r);
        const uint32_t e = table[hring_peek(r, max_bits)];
        out[n] = (uint8_t)e;
        hring_skip(r, e >> 8);
        clean |= r.remaining == 0;
    }
    hring_fill(r);
    if (max_bits >= 1) {
        // 16 symbols at a time.  Per symbol: the cell's index is the top max_bits bits of the window's high word (one shift: a
        // refill leaves at least 32 valid bits there, and max_bits <= 11), one 16-bit load at a 32-bit shared-memory address, the
        // skip.  Whether the bits run out exactly at a symbol boundary (`clean`) can only happen in a group that starts with
        // less than 16 x max_bits bits left: only those groups look after every symbol.
        const uint32_t tb = (uint32_t)__cvta_generic_to_shared(table);
        const uint32_t psh = 32 - max_bits;
        const int32_t near_end = (int32_t)(16 * max_bits);
        while (n + 16 <= expected && r.remaining > near_end) {
            huf_group16<false>(r, tb, psh, out + n, clean);
            n += 16;
        }
        while (n + 16 <= expected) {
            huf_group16<true>(r, tb, psh, out + n, clean);
            n += 16;
        }
    }
    hring_fill(r);
    for (; n < expected; n++) {
        hring_refill(r);
        const uint32_t e = table[hring_peek(r, max_bits)];
        out[n] = (uint8_t)e;
        hring_skip(r, e >> 8);
        clean |= r.remaining == 0;
    }
    if (r.remaining > 0) return SZB_ERR_PANIC;  // more symbols than its slot holds: the reference indexes past its output slice
    // ended early but on a symbol boundary: the caller of a 4-stream block compares the count (literals.go:320,332,349); for a
    // single stream the reference does not (it goes on with what its reused buffer held): the engine's documented -19
    if (r.remaining < 0) return clean && one_of_four ? SZB_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH : SZB_ERR_DIDNT_USE_ALL_BITS_TO_DECODE_HUFFMAN;
    return SZB_OK;
}

constexpr uint32_t kHufGroup = 8;         // blocks per warp in k_decode_literals (4 lanes each)
constexpr uint32_t kHufCellsPerWarp = 2048;  // 4 KB of decode tables resident per warp (one maxBits-11 table)

// Stage 2: one LANE per Huffman stream (literals.go:295-371): a warp decodes the streams of up to 8
// blocks at once.  The blocks' decode tables are copied from the arena into the warp's 8 KB of shared
// memory; neighbouring blocks that share a table (a Compressed block followed by its Treeless users)
// share one copy.  When the tables do not fit (maxBits 11 = 4 KB each) the group is done in passes.
// 6.5 KB of shared memory per warp (tables + bit rings): eight CTAs per SM; a deeper ring (three copy
// groups in flight) costs two of them and measured slower.
__global__ void __launch_bounds__(kCtaThreads) k_decode_literals(DeviceBatch a) {
    __shared__ __align__(16) uint16_t tabs_all[kWarpsPerCta][kHufCellsPerWarp];
    __shared__ __align__(16) uint8_t rings_all[kWarpsPerCta][32 * kHufRingStride];
    const uint32_t warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = blockIdx.x * kWarpsPerCta + warp_in_cta;
    const uint32_t first = g * kHufGroup;
    if (first >= a.n_huf) return;
    uint16_t *tabs = tabs_all[warp_in_cta];
    uint8_t *my_ring = rings_all[warp_in_cta] + lane * kHufRingStride;
    const uint32_t n_entries = a.n_huf - first < kHufGroup ? a.n_huf - first : kHufGroup;

    // lanes 0..7 fetch the per-entry facts; everyone reads them through shuffles
    uint32_t e_slot = 0, e_bits = 0, e_tree = 0;
    int e_rc = SZB_OK;
    if (lane < n_entries) {
        e_slot = a.huf_slot[first + lane];
        const HufInfo info = a.huf_info[e_slot];
        e_bits = info.max_bits;
        e_tree = info.tree_bytes;
        e_rc = info.status;
    }
    const uint32_t my_e = lane >> 2, my_k = lane & 3;
    int my_rc = SZB_OK;
    const int rc0 = __shfl_sync(kFull, e_rc, my_e);
    const uint32_t bits = __shfl_sync(kFull, e_bits, my_e);
    const uint32_t tree_b = __shfl_sync(kFull, e_tree, my_e);

    uint32_t next_e = 0;
    while (next_e < n_entries) {
        // take entries while their tables fit; consecutive entries with the same table share it
        uint32_t used = 0, count = 0, my_off = 0;
        uint32_t prev_slot = 0xFFFFFFFFu, prev_off = 0;
        for (uint32_t e = next_e; e < n_entries; e++) {
            const uint32_t slot = __shfl_sync(kFull, e_slot, e);
            const uint32_t ebits = __shfl_sync(kFull, e_bits, e);
            const int rc = __shfl_sync(kFull, e_rc, e);
            uint32_t off;
            if (rc != SZB_OK) {
                off = 0;  // nothing to load: the block inherits the failure
            } else if (slot == prev_slot) {
                off = prev_off;
            } else {
                const uint32_t size = 1u << ebits;
                if (used + size > kHufCellsPerWarp) break;
                off = used;
                const uint16_t *src_tab = a.huf_tabs + (size_t)slot * (1u << kMaxHufBits);
                for (uint32_t i = lane; i < size; i += 32) tabs[off + i] = src_tab[i];
                used += size;
                prev_slot = slot;
                prev_off = off;
            }
            if (e == my_e) my_off = off;
            count++;
        }
        __syncwarp();
        if (my_e >= next_e && my_e < next_e + count) {
            const uint32_t b = a.huf_list[first + my_e];
            const szb_block_desc d = a.blocks[b];
            if (rc0 != SZB_OK) {
                my_rc = rc0;
            } else if (d.flags & SZB_BLOCK_TABLES_ONLY) {
                my_rc = SZB_OK;  // a dictionary's row: its tree was built, it has no literals
            } else {
                // --- this block's streams (literals.go:270-371) ---
                const uint8_t *payload = a.src + d.src_off;
                const uint32_t own_tree = d.lit_type == 2 ? tree_b : 0;
                const uint32_t skip = d.lit_hdr_bytes + own_tree;
                int32_t comp = (int32_t)d.lit_comp - (int32_t)own_tree;
                uint8_t *out = a.litbuf + d.lit_buf_off;
                const uint32_t regen = d.lit_regen;
                if (d.lit_streams == 1) {
                    if (comp < 0)
                        my_rc = SZB_ERR_PANIC;
                    else if (my_k == 0)
                        my_rc = huf_decode_stream_vec(tabs + my_off, bits, payload + skip, (uint32_t)comp, out, regen, my_ring, false);
                } else {
                    comp -= 6;
                    if (comp < 0) {
                        my_rc = SZB_ERR_PANIC;  // literals.go:283 negative slice bound
                    } else {
                        const uint8_t *jt = payload + skip;  // literals.go:46-58 jump table, 3 x u16 LE
                        const uint32_t s1 = jt[0] | (jt[1] << 8), s2 = jt[2] | (jt[3] << 8), s3 = jt[4] | (jt[5] << 8);
                        const uint32_t normal = (regen + 3) / 4;  // literals.go:306-311
                        const int32_t last = (int32_t)regen - 3 * (int32_t)normal;
                        // The reference decodes the streams one after the other and checks every stream's extent right before it
                        // decodes it (literals.go:313-359): stream k's extent error comes AFTER the decode errors of the streams
                        // before it.  Every lane applies the checks of its own stream; the first failing stream decides (below).
                        const uint32_t end1 = s1, end2 = s1 + s2, end3 = s1 + s2 + s3;
                        const uint32_t s4 = (uint32_t)(uint16_t)((uint32_t)comp - (uint32_t)(uint16_t)end3);  // CalcStreamsize4: uint16 arithmetic
                        const uint32_t my_end = my_k == 0 ? end1 : (my_k == 1 ? end2 : (my_k == 2 ? end3 : end3 + s4));
                        const uint32_t before = my_k == 0 ? 0 : (my_k == 1 ? end1 : (my_k == 2 ? end2 : end3));
                        if (last < 0) {
                            my_rc = SZB_ERR_PANIC;  // literals.go:311 inverted slice
                        } else if (before > (uint32_t)comp) {
                            my_rc = SZB_ERR_CORRUPTED_JUMPTABLE;  // an earlier stream reports it first; never decoded from here
                        } else if (my_k == 2 && my_end > regen) {
                            my_rc = SZB_ERR_PANIC;  // literals.go:339-342 panic("Corrupt stream sizes")
                        } else if (my_k < 3 && my_end > (uint32_t)comp) {
                            my_rc = SZB_ERR_CORRUPTED_JUMPTABLE;  // literals.go:54-56 (the reference re-slices into stale capacity)
                        } else if (my_k == 3 && my_end != (uint32_t)comp) {
                            my_rc = SZB_ERR_PANIC;  // literals.go:356-359
                        } else {
                            const uint32_t expected = my_k < 3 ? normal : (uint32_t)last;
                            my_rc = huf_decode_stream_vec(tabs + my_off, bits, jt + 6 + before, my_end - before, out + my_k * normal, expected, my_ring, true);
                            // the fourth stream's count is only checked through the sum of all four: a panic (literals.go:366-369)
                            if (my_k == 3 && my_rc == SZB_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH) my_rc = SZB_ERR_PANIC;
                        }
                    }
                }
            }
        }
        __syncwarp();
        next_e += count;
    }
    // per block: the first failing stream in the order 1..4 decides (the reference decodes them in order)
    int rc = my_rc;
    for (int k = 1; k < 4; k++) {
        const int other = __shfl_sync(kFull, my_rc, (lane & ~3u) + k);
        if (rc == SZB_OK) rc = other;
    }
    if (my_k == 0 && my_e < n_entries) a.lit_status[a.huf_list[first + my_e]] = rc;
}

// resident cell: symbol in bits 0-5, next-state counter in bits 6-15
__device__ __forceinline__ uint32_t cell16(uint32_t packed, uint32_t al) {
    return fse_code(packed) | (((fse_baseline(packed) + (1u << al)) >> fse_nb(packed)) << 6);
}

// shared memory per warp, k_build_seq_tables
struct SeqSmem {
    uint32_t tll[1 << kMaxALLL];
    uint32_t tml[1 << kMaxALML];
    uint32_t tof[1 << kMaxALOF];
    uint8_t symk[1 << kMaxALLL];
    int16_t norm[kMaxFseSymbols];
    uint16_t next[kMaxFseSymbols];
};

// Stage 1 for the sequences section: one warp per compressed block that has sequences builds
// the block's LL / OF / ML decode tables in shared memory (DecodeTables, sequences.go:275-369)
// and publishes them to the table arena in HBM (slot = position in seq_list), together with
// the accuracy logs and the offset of the backward bitstream.
__global__ void __launch_bounds__(kCtaThreads) k_build_seq_tables(DeviceBatch a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kWarpsPerCta + warp_in_cta;
    if (w >= a.n_seq) return;
    SeqSmem &sm = reinterpret_cast<SeqSmem *>(smem_raw)[warp_in_cta];
    const uint32_t b = a.seq_list[w];
    const szb_block_desc d = a.blocks[b];
    const uint8_t *tables = a.src + d.src_off + d.seq_off + d.seq_hdr_bytes;
    const uint32_t tables_avail = d.block_size - d.seq_off - d.seq_hdr_bytes;

    uint32_t cursor = 0;
    uint32_t al[3] = {0, 0, 0};
    int rc = SZB_OK;
#pragma unroll
    for (int i = 0; i < 3 && rc == SZB_OK; i++) {
        const int kind = i;  // KIND_LL, KIND_OF, KIND_ML: the order of the table bytes
        uint32_t *table = kind == KIND_LL ? sm.tll : (kind == KIND_OF ? sm.tof : sm.tml);
        const uint32_t mode = field_mode(d.seq_modes, kind);
        TableSource ts;
        if (mode == 3) {  // Repeat: rebuild from the origin block's bytes (the host chased the chain)
            const uint32_t ob = kind == KIND_LL ? d.ll_origin : (kind == KIND_OF ? d.of_origin : d.ml_origin);
            const szb_block_desc o = a.blocks[ob];
            int lrc = SZB_OK;
            if (lane == 0)
                lrc = locate_field(a.src + o.src_off + o.seq_off + o.seq_hdr_bytes, o.block_size - o.seq_off - o.seq_hdr_bytes,
                                   o.seq_modes, kind, sm.norm, &ts);
            rc = __shfl_sync(kFull, lrc, 0);
            ts.p = (const uint8_t *)__shfl_sync(kFull, (unsigned long long)ts.p, 0);
            ts.avail = __shfl_sync(kFull, ts.avail, 0);
            ts.mode = __shfl_sync(kFull, ts.mode, 0);
            if (rc) break;
        } else {
            ts.p = tables + cursor;
            ts.avail = tables_avail - cursor;
            ts.mode = mode;
        }
        uint32_t used = 0;
        rc = build_fse_table(ts, kind, a.predef, table, sm.norm, sm.next, sm.symk, &al[i], &used);
        if (mode != 3) cursor += used;
        if (rc == SZB_OK && cursor > tables_avail) rc = SZB_ERR_UNEXPECTED_EOF;
    }
    if (rc != SZB_OK) {
        if (lane == 0) a.seq_status[b] = rc;
        return;
    }
    // published as the 16-bit resident cells k_decode_sequences works on (symbol | next << 6)
    uint16_t *slot = a.seq_tabs + (size_t)w * kTabSlotWords;
    for (uint32_t i = lane; i < (1u << al[KIND_LL]); i += 32) slot[i] = (uint16_t)cell16(sm.tll[i], al[KIND_LL]);
    for (uint32_t i = lane; i < (1u << al[KIND_ML]); i += 32) slot[512 + i] = (uint16_t)cell16(sm.tml[i], al[KIND_ML]);
    for (uint32_t i = lane; i < (1u << al[KIND_OF]); i += 32) slot[1024 + i] = (uint16_t)cell16(sm.tof[i], al[KIND_OF]);
    if (lane == 0) {
        SeqInfo info;
        info.al_ll = (uint8_t)al[KIND_LL];
        info.al_of = (uint8_t)al[KIND_OF];
        info.al_ml = (uint8_t)al[KIND_ML];
        info.pad = 0;
        info.stream_off = cursor;
        a.seq_info[w] = info;
    }
}

// Stage 3: DecodeSequences (sequences.go:126-206).  One LANE per block: a warp decodes
// kSeqLanes blocks in lock step, each lane walking its own backward bitstream with its own three
// FSE states, the tables of all its blocks resident in shared memory.  A warp instruction thus
// advances kSeqLanes independent state chains instead of one.
//
// Shared-memory budget decides how many chains an SM can run, so the resident cells are 16 bits:
// symbol (6 bits) and the cell's "next state" counter (10 bits, fse.go:195-196), from which
// NumberOfBits = AL - highbit(next) and Baseline = (next << NumberOfBits) - 2^AL (fse.go:212-213)
// are recomputed with two ALU ops each; extra-bit counts and base values (predefined.go:5-20,36-50)
// come from a 64-entry table shared by the warp.
//
// Bit reads (replacing Reversebitstream.Read, reversebitstream.go:17-88): every lane keeps the
// 128 stream bytes around its read position in an 8 x 16-byte shared-memory ring, topped up with
// 16-byte cp.async copies once per group of four sequences, two groups ahead of use.  Per sequence
// a lane builds one 64-bit window ending at its bit position from three ring words and peels the
// six fields off its top in the reference's order: OF extra, ML extra, LL extra, then (except after
// the last sequence) LL state, ML state, OF state.
//
// The four sequences of a group are decoded speculatively on that branch-free path; a sequence
// wider than the window, a read below the start of the stream or a ring that is not far enough ahead
// only sets a flag, and a flagged group is decoded again from its saved state with byte reads
// (slow_step).  Decoded triples leave as 16-byte stores, one group at a time.
constexpr uint32_t kSeqTabBytes = kSeqLanes * kTabSlotWords * 2;  // u16 cells
constexpr uint32_t kSeqRingBytes = 128;
constexpr uint32_t kSeqRingStride = kSeqRingBytes + 16;           // per lane: the ring + a mirror of its first 16 B
constexpr uint32_t kSeqRingOff = kSeqTabBytes;                    // byte offset of the rings
constexpr uint32_t kSeqLutWord = (kSeqRingOff + kSeqLanes * kSeqRingStride) / 4;  // ll[64] | ml[64]: base | extra << 24
constexpr uint32_t kSeqBarWord = kSeqLutWord + 128;                // the table load's mbarrier (8 bytes, 8-byte aligned)
static_assert(kSeqBarWord % 2 == 0, "the mbarrier is a 64-bit object");
constexpr uint32_t kSeqDecodeSmemBytes = (kSeqBarWord + 2) * 4;
// A group of four window reads consumes at most 4 x 57 bits = 28.5 bytes and a window reaches 10
// bytes below the byte of its top bit: a group touches nothing below (top - kSeqGroupReach).
// The top-up of group g requests chunks down to the one holding (top - kSeqRingAhead); what it
// requests is only waited for at group g+2 (wait_group 2), whose top is at most 57 bytes lower:
// kSeqRingAhead >= 57 + kSeqGroupReach.  The ring then spans at most chunk(top - 96) .. chunk(top + 4),
// and a copy for chunk c lands in the slot of chunk c + 8 >= chunk(top + 32): never one still being read.
constexpr int32_t kSeqGroupReach = 39;
constexpr int32_t kSeqRingAhead = 96;

__device__ __forceinline__ uint32_t bfind(uint32_t x) {  // index of the highest set bit (x != 0)
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}

struct SeqLane {
    uint32_t s_ll, s_of, s_ml;  // FSE states
    int32_t pos;                // stream bits not consumed yet
};

// Chunk c (16 B, counted from the aligned chunk holding sp[0]) goes to ring slot c & 7, slot 0 also
// to the mirror above slot 7 so that a 12-byte window read never wraps.  cp.async: no register and
// no scoreboard slot is held while the copy flies; the caller commits and waits.
__device__ __forceinline__ void ring_fetch_if(bool p, uint32_t ring_saddr, const uint4 *chunk0, int32_t c) {
    const uint32_t slot = (uint32_t)c & 7;
    const uint32_t dst = ring_saddr + (slot << 4);
    const uint4 *src = chunk0 + c;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %0, 0;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@p cp.async.cg.shared.global [%2], [%4], 16;\n\t"
        "@q cp.async.cg.shared.global [%3], [%4], 16;\n\t}"
        ::"r"((uint32_t)p), "r"((uint32_t)(p && slot == 0)), "r"(dst), "r"(ring_saddr + kSeqRingBytes), "l"(src)
        : "memory");
}
// Stream start: requests every chunk below `lowreq` down to the one holding byte (top - kSeqRingAhead).
__device__ __forceinline__ void ring_fill(uint32_t ring_saddr, const uint4 *chunk0, int32_t top, int32_t &lowreq) {
    int32_t need = (top - kSeqRingAhead) >> 4;
    need = need < 0 ? 0 : need;
    while (lowreq > need) ring_fetch_if(true, ring_saddr, chunk0, --lowreq);
}
// Steady state, no branch: at most two more chunks, which is more than a group of window reads consumes.
__device__ __forceinline__ void ring_topup(uint32_t ring_saddr, const uint4 *chunk0, int32_t top, int32_t &lowreq) {
    int32_t need = (top - kSeqRingAhead) >> 4;
    need = need < 0 ? 0 : need;
    const bool p1 = lowreq - 1 >= need, p2 = lowreq - 2 >= need;
    ring_fetch_if(p1, ring_saddr, chunk0, lowreq - 1);
    ring_fetch_if(p2, ring_saddr, chunk0, lowreq - 2);
    lowreq -= (int32_t)p1 + (int32_t)p2;
}

// One sequence on the window path.  `bad` goes negative when the result must not be used: more than
// 57 bits wanted, or more than the stream still holds (the zero fill of reversebitstream.go:67-75
// is left to slow_step).  Everything it touches stays in bounds whatever the bits are: ring offsets
// are masked and an n-bit field cannot push a state out of its table.
template <bool kUpdate>
__device__ __forceinline__ void fast_step(const uint32_t *sw, const uint16_t *tll, const uint16_t *tml, const uint16_t *tof,
                                          const uint8_t *ring, int32_t sp_mis, uint32_t al_ll, uint32_t al_ml, uint32_t al_of,
                                          SeqLane &L, int32_t &bad, uint32_t &v_ll, uint32_t &v_ml, uint32_t &v_of) {
    const uint32_t c_of = tof[L.s_of], c_ll = tll[L.s_ll], c_ml = tml[L.s_ml];  // peek OF, LL, ML (sequences.go:67-78)
    const uint32_t ofc = c_of & 63;
    const uint32_t u_ll = sw[kSeqLutWord + (c_ll & 63)], u_ml = sw[kSeqLutWord + 64 + (c_ml & 63)];
    const uint32_t llx = u_ll >> 24, mlx = u_ml >> 24;
    uint32_t nbl = 0, nbm = 0, nbo = 0;
    if (kUpdate) {  // NumberOfBits = AL - highbit(next) (fse.go:212)
        nbl = al_ll - bfind(c_ll >> 6);
        nbm = al_ml - bfind(c_ml >> 6);
        nbo = al_of - bfind(c_of >> 6);
    }
    const uint32_t total = ofc + mlx + llx + nbl + nbm + nbo;
    // window: the 8 bytes ending at the byte that holds bit pos-1 (byte `top`, counted from chunk 0),
    // that bit moved to bit 63
    const int32_t top = sp_mis + ((L.pos - 1) >> 3);
    const uint32_t ap = (uint32_t)(top - 7);
    const uint32_t mis = ap & 3;
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(ring + ((ap - mis) & (kSeqRingBytes - 1)));
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    uint32_t lo = __funnelshift_r(w0, w1, mis * 8), hi = __funnelshift_r(w1, w2, mis * 8);
    const uint32_t k = 7 - ((uint32_t)(L.pos - 1) & 7);
    hi = __funnelshift_l(lo, hi, k);
    lo <<= k;
    uint32_t x_of, x_ml, x_ll, b_ll = 0, b_ml = 0, b_of = 0;
#define SZB_TAKE(dst, n)                \
    dst = __funnelshift_l(hi, 0u, (n)); \
    hi = __funnelshift_l(lo, hi, (n));  \
    lo <<= (n);
    SZB_TAKE(x_of, ofc)
    SZB_TAKE(x_ml, mlx)
    SZB_TAKE(x_ll, llx)
    if (kUpdate) {
        SZB_TAKE(b_ll, nbl)
        SZB_TAKE(b_ml, nbm)
        SZB_TAKE(b_of, nbo)
    }
#undef SZB_TAKE
    L.pos -= (int32_t)total;
    bad |= (57 - (int32_t)total) | L.pos;
    v_of = (1u << ofc) + x_of;              // sequences.go:99-104
    v_ml = (u_ml & 0xFFFFFF) + x_ml;        // sequences.go:106-112
    v_ll = (u_ll & 0xFFFFFF) + x_ll;        // sequences.go:114-120
    if (kUpdate) {  // update LL, ML, OF (sequences.go:178-194); Baseline = (next << nb) - 2^AL (fse.go:213)
        L.s_ll = ((c_ll >> 6) << nbl) - (1u << al_ll) + b_ll;
        L.s_ml = ((c_ml >> 6) << nbm) - (1u << al_ml) + b_ml;
        L.s_of = ((c_of >> 6) << nbo) - (1u << al_of) + b_of;
    }
}

// n (<= 32) bits ending at bit position pos of a backward stream, i.e. bits [pos-n, pos) of the
// little-endian integer; zero below bit 0 (reversebitstream.go:23-27,67-75).  Byte loads.
__device__ __noinline__ uint32_t slow_read_bits(const uint8_t *sp, int32_t pos, uint32_t n) {
    if (n == 0 || pos <= 0) return 0;
    const int32_t lo = pos - (int32_t)n;
    const int32_t l = lo < 0 ? 0 : lo;
    const int32_t first = l >> 3, last = (pos - 1) >> 3;
    uint64_t acc = 0;
    for (int32_t bb = first; bb <= last; bb++) acc |= (uint64_t)sp[bb] << (8 * (bb - first));
    acc >>= (l & 7);
    acc &= (1ull << (uint32_t)(pos - l)) - 1;
    if (lo < 0) acc <<= (uint32_t)(-lo);
    return (uint32_t)acc;
}

// The same sequence with one byte-wise read per field: any width, any position.  Rare.
template <bool kUpdate>
__device__ __noinline__ void slow_step(const uint32_t *sw, const uint16_t *tll, const uint16_t *tml, const uint16_t *tof,
                                       const uint8_t *sp, uint32_t al_ll, uint32_t al_ml, uint32_t al_of, SeqLane &L,
                                       uint32_t &v_ll, uint32_t &v_ml, uint32_t &v_of) {
    const uint32_t c_of = tof[L.s_of], c_ll = tll[L.s_ll], c_ml = tml[L.s_ml];
    const uint32_t ofc = c_of & 63;
    const uint32_t u_ll = sw[kSeqLutWord + (c_ll & 63)], u_ml = sw[kSeqLutWord + 64 + (c_ml & 63)];
    const uint32_t llx = u_ll >> 24, mlx = u_ml >> 24;
    const uint32_t x_of = slow_read_bits(sp, L.pos, ofc);
    L.pos -= (int32_t)ofc;
    const uint32_t x_ml = slow_read_bits(sp, L.pos, mlx);
    L.pos -= (int32_t)mlx;
    const uint32_t x_ll = slow_read_bits(sp, L.pos, llx);
    L.pos -= (int32_t)llx;
    v_of = (1u << ofc) + x_of;
    v_ml = (u_ml & 0xFFFFFF) + x_ml;
    v_ll = (u_ll & 0xFFFFFF) + x_ll;
    if (kUpdate) {
        const uint32_t nbl = al_ll - bfind(c_ll >> 6), nbm = al_ml - bfind(c_ml >> 6), nbo = al_of - bfind(c_of >> 6);
        const uint32_t b_ll = slow_read_bits(sp, L.pos, nbl);
        L.pos -= (int32_t)nbl;
        const uint32_t b_ml = slow_read_bits(sp, L.pos, nbm);
        L.pos -= (int32_t)nbm;
        const uint32_t b_of = slow_read_bits(sp, L.pos, nbo);
        L.pos -= (int32_t)nbo;
        L.s_ll = ((c_ll >> 6) << nbl) - (1u << al_ll) + b_ll;
        L.s_ml = ((c_ml >> 6) << nbm) - (1u << al_ml) + b_ml;
        L.s_of = ((c_of >> 6) << nbo) - (1u << al_of) + b_of;
    }
}

// One group of kSeqLanes blocks on one warp.  `use` counts the groups this CTA has decoded: the table load's mbarrier is
// initialised once per CTA and completes one phase per group.
__device__ __forceinline__ void seq_decode_group(const DeviceBatch &a, uint32_t *sw, uint32_t first, uint32_t use, uint32_t lane) {
    const uint32_t n_here = a.n_seq - first < (uint32_t)kSeqLanes ? a.n_seq - first : (uint32_t)kSeqLanes;
    uint16_t *tabs = reinterpret_cast<uint16_t *>(sw);

    // tables: HBM arena -> shared memory, 2560 contiguous bytes per block.  One bulk copy per block (cp.async.bulk, the copy engine
    // of sm_90+: one instruction per lane instead of 160 16-byte cp.async's), all in flight at once, their bytes counted by an
    // mbarrier the warp then waits on.
    {
        const uint32_t tabs_saddr = (uint32_t)__cvta_generic_to_shared(tabs);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(sw + kSeqBarWord);
        const uint8_t *arena = reinterpret_cast<const uint8_t *>(a.seq_tabs + (size_t)first * kTabSlotWords);
        constexpr uint32_t kSlotBytes = kTabSlotWords * 2;
        if (lane == 0) {
            if (use == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            // the barrier's initialisation, and the previous group's reads of the cells (generic proxy), come before what the copy
            // engine does to both (async proxy)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n_here * kSlotBytes) : "memory");
        __syncwarp();
        if (lane < n_here)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tabs_saddr + lane * kSlotBytes),
                         "l"(arena + (size_t)lane * kSlotBytes), "r"(kSlotBytes), "r"(bar)
                         : "memory");
        if (use == 0)
            for (uint32_t i = lane; i < 64; i += 32) {
                sw[kSeqLutWord + i] = kLLBaseDev[i] | ((uint32_t)kLLExtraDev[i] << 24);
                sw[kSeqLutWord + 64 + i] = kMLBaseDev[i] | ((uint32_t)kMLExtraDev[i] << 24);
            }
        uint32_t landed = 0;
        while (!landed)  // the phase completes when the expected bytes have arrived (try_wait sleeps in hardware, it does not spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(landed) : "r"(bar), "r"(use & 1) : "memory");
    }
    __syncwarp();
    if (lane >= n_here) return;

    const uint32_t w = first + lane;
    const uint32_t b = a.seq_list[w];
    if (a.seq_status[b] != SZB_OK) return;  // its tables failed to build
    const szb_block_desc d = a.blocks[b];
    if (d.flags & SZB_BLOCK_TABLES_ONLY) {  // a dictionary's row: its tables were built, it has no sequences to decode
        a.out_size[b] = 0;
        return;
    }
    const SeqInfo info = a.seq_info[w];
    const uint16_t *tll = tabs + lane * kTabSlotWords, *tml = tll + 512, *tof = tll + 1024;
    const uint32_t al_ll = info.al_ll, al_ml = info.al_ml, al_of = info.al_of;
    const uint32_t hdr = d.seq_off + d.seq_hdr_bytes + info.stream_off;
    const uint8_t *sp = a.src + d.src_off + hdr;
    const uint32_t len = d.block_size - hdr;

    // padding: zero bits then the first 1 bit, at most 8 (sequences.go:131-143)
    if (len == 0 || sp[len - 1] == 0) {
        a.seq_status[b] = SZB_ERR_BAD_PADDING;
        return;
    }
    SeqLane L;
    L.pos = (int32_t)(len * 8) - (__clz((uint32_t)sp[len - 1]) - 24 + 1);
    // InitState in the order LL, OF, ML (sequences.go:145-159)
    L.s_ll = slow_read_bits(sp, L.pos, al_ll);
    L.pos -= (int32_t)al_ll;
    L.s_of = slow_read_bits(sp, L.pos, al_of);
    L.pos -= (int32_t)al_of;
    L.s_ml = slow_read_bits(sp, L.pos, al_ml);
    L.pos -= (int32_t)al_ml;

    const uint8_t *ring = reinterpret_cast<const uint8_t *>(sw) + kSeqRingOff + lane * kSeqRingStride;
    const uint32_t ring_saddr = (uint32_t)__cvta_generic_to_shared(ring);
    const int32_t sp_mis = (int32_t)(reinterpret_cast<uintptr_t>(sp) & 15);
    const uint4 *chunk0 = reinterpret_cast<const uint4 *>(sp - sp_mis);
    // lowest chunk requested so far; everything from the chunk of the last stream byte down is wanted
    int32_t lowreq = ((sp_mis + (int32_t)len - 1) >> 4) + 1;
    ring_fill(ring_saddr, chunk0, sp_mis + ((L.pos - 1) >> 3), lowreq);
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    const uint32_t nseq = d.nseq;
    uint32_t *gll = a.seq_ll + d.seq_buf_off, *gml = a.seq_ml + d.seq_buf_off, *gof = a.seq_of + d.seq_buf_off;
    uint64_t ml_sum = 0;
    const uint32_t n_upd = nseq - 1;  // every sequence but the last updates the states (sequences.go:178)
    uint32_t i = 0;
    int32_t low1 = lowreq, low2 = lowreq;  // lowreq one and two top-ups ago
    for (; i + 4 <= n_upd; i += 4) {
        // what was requested two top-ups ago has landed once wait_group 2 returns
        const int32_t landed = low2 ? (low2 << 4) : -64;
        const int32_t top = sp_mis + ((L.pos - 1) >> 3);
        ring_topup(ring_saddr, chunk0, top, lowreq);
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        low2 = low1;
        low1 = lowreq;
        const SeqLane S = L;
        int32_t bad = top - kSeqGroupReach - landed;
        uint32_t v_ll[4], v_ml[4], v_of[4];
#pragma unroll
        for (int j = 0; j < 4; j++) fast_step<true>(sw, tll, tml, tof, ring, sp_mis, al_ll, al_ml, al_of, L, bad, v_ll[j], v_ml[j], v_of[j]);
        if (bad < 0) {  // T and o are what the out-of-line call sees: L and the v arrays stay in registers
            SeqLane T = S;
            uint32_t o[12];
            for (int j = 0; j < 4; j++) slow_step<true>(sw, tll, tml, tof, sp, al_ll, al_ml, al_of, T, o[j], o[4 + j], o[8 + j]);
            L = T;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                v_ll[j] = o[j];
                v_ml[j] = o[4 + j];
                v_of[j] = o[8 + j];
            }
        }
        // seq_buf_off is a multiple of 32 entries: 16-byte aligned stores
        *reinterpret_cast<uint4 *>(gll + i) = make_uint4(v_ll[0], v_ll[1], v_ll[2], v_ll[3]);
        *reinterpret_cast<uint4 *>(gml + i) = make_uint4(v_ml[0], v_ml[1], v_ml[2], v_ml[3]);
        *reinterpret_cast<uint4 *>(gof + i) = make_uint4(v_of[0], v_of[1], v_of[2], v_of[3]);
        ml_sum += (uint64_t)v_ml[0] + v_ml[1] + v_ml[2] + v_ml[3];
    }
    {   // the last (at most four) sequences one at a time; the ring is made to cover them all at once
        ring_fill(ring_saddr, chunk0, sp_mis + ((L.pos - 1) >> 3), lowreq);
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const int32_t landed = lowreq ? (lowreq << 4) : -64;
        for (; i < nseq; i++) {
            uint32_t v_ll, v_ml, v_of;
            const SeqLane S = L;
            int32_t bad = sp_mis + ((L.pos - 1) >> 3) - 10 - landed;
            if (i < n_upd)
                fast_step<true>(sw, tll, tml, tof, ring, sp_mis, al_ll, al_ml, al_of, L, bad, v_ll, v_ml, v_of);
            else
                fast_step<false>(sw, tll, tml, tof, ring, sp_mis, al_ll, al_ml, al_of, L, bad, v_ll, v_ml, v_of);
            if (bad < 0) {
                SeqLane T = S;
                uint32_t o[3];
                if (i < n_upd)
                    slow_step<true>(sw, tll, tml, tof, sp, al_ll, al_ml, al_of, T, o[0], o[1], o[2]);
                else
                    slow_step<false>(sw, tll, tml, tof, sp, al_ll, al_ml, al_of, T, o[0], o[1], o[2]);
                L = T;
                v_ll = o[0];
                v_ml = o[1];
                v_of = o[2];
            }
            gll[i] = v_ll;
            gml[i] = v_ml;
            gof[i] = v_of;
            ml_sum += v_ml;
        }
    }
    // the stream must be consumed exactly (sequences.go:197-204)
    a.seq_status[b] = L.pos == 0 ? SZB_OK : SZB_ERR_NOT_ALL_BITS_USED;
    a.out_size[b] = (uint64_t)d.lit_regen + ml_sum;
}

// One group per CTA: the default launch.
__global__ void __launch_bounds__(32) k_decode_sequences(DeviceBatch a) {
    extern __shared__ __align__(16) uint32_t sw[];
    seq_decode_group(a, sw, blockIdx.x * kSeqLanes, 0, threadIdx.x);
}

// The same with a grid smaller than the number of groups (launch_entropy: SZB_SEQ_CTAS_PER_SM, a cap on the CTAs an SM holds --
// experiments with other kernels beside this one, and tests): a CTA walks the groups blockIdx.x, + gridDim.x, ...
__global__ void __launch_bounds__(32) k_decode_sequences_multi(DeviceBatch a) {
    extern __shared__ __align__(16) uint32_t sw[];
    const uint32_t lane = threadIdx.x;
    const uint32_t ngroups = (a.n_seq + kSeqLanes - 1) / kSeqLanes;
    uint32_t use = 0;
    for (uint32_t g = blockIdx.x; g < ngroups; g += gridDim.x, use++) {
        seq_decode_group(a, sw, g * kSeqLanes, use, lane);
        __syncwarp();  // every lane is done with the group's cells and rings
    }
}

#include "sequences3.cuh"
#include "execute.cuh"

}  // namespace szb
