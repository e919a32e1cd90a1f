// kernels.cuh -- the four GPU stages of the decode path (sm_100a).
//
//   k_huffman_literals  stage 1+2: Huffman tree description (direct / FSE-compressed weights),
//                       decode table in shared memory, 1- or 4-stream literal decode
//                       (structure/huffman.go, structure/literals.go:209-373)
//   k_sequences         stage 1+3: LL/OF/ML table selection + construction in shared memory,
//                       three interleaved FSE states over the backward bitstream
//                       (fse/fse.go, fse/predefined.go, structure/sequences.go)
//   k_scan_blocks       device-wide exclusive prefix sum of per-block regenerated sizes
//                       (the reference gets positions for free from its ring buffer,
//                       decompression/ringbuffer.go:102-178)
//   k_execute           stage 4: literal copies, repeat-offset history, match copies against
//                       the output itself, Raw / RLE block bodies
//                       (decompression/sequence_execution.go:14-114, ringbuffer.go:197-277,
//                       framedecompressor.go:211-241)
//
// All arithmetic is integer; there is no tensor-core work on this path.
#pragma once
#include <cuda_runtime.h>

#include "huffman.cuh"
#include "sequences.cuh"

namespace szb {

constexpr int kWarpsPerCta = 4;
constexpr int kCtaThreads = kWarpsPerCta * 32;
constexpr uint32_t kFull = 0xFFFFFFFFu;

#ifndef SZB_SERIAL_TABLES
#define SZB_SERIAL_TABLES 0
#endif

struct DeviceBatch {
    const uint8_t *src;
    const szb_block_desc *blocks;
    const szb_frame_desc *frames;
    uint32_t nblocks, nframes;
    const uint32_t *huf_list;  // blocks with Huffman-coded literals
    uint32_t n_huf;
    const uint32_t *seq_list;  // blocks with nseq > 0
    uint32_t n_seq;
    uint8_t *litbuf;
    uint32_t *seq_ll, *seq_ml, *seq_of;
    uint64_t *out_size;     // per block regenerated size (host-initialised for Raw/RLE/zero-sequence blocks)
    uint64_t *out_off;      // per block exclusive prefix
    int32_t *lit_status;    // per block
    int32_t *seq_status;    // per block
    uint64_t *total;        // [0] = total output bytes
    const uint32_t *predef; // predefined LL(64) | OF(32) | ML(64) decode tables
    uint8_t *dst;
    uint64_t dst_cap;
    uint64_t *frame_out_off, *frame_out_len;
    int32_t *frame_status;
};

__device__ __forceinline__ int warp_first_error(int rc) {
    uint32_t bad = __ballot_sync(kFull, rc != 0);
    if (!bad) return 0;
    return __shfl_sync(kFull, rc, __ffs(bad) - 1);
}

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

// shared memory per warp, k_huffman_literals
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

// One warp per block with Huffman-coded (Compressed or Treeless) literals.
__global__ void __launch_bounds__(kCtaThreads) k_huffman_literals(DeviceBatch a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kWarpsPerCta + warp_in_cta;
    if (w >= a.n_huf) return;
    HufSmem &sm = reinterpret_cast<HufSmem *>(smem_raw)[warp_in_cta];
    const uint32_t b = a.huf_list[w];
    const szb_block_desc d = a.blocks[b];
    const szb_block_desc o = a.blocks[d.huf_origin];

    // --- tree description of the origin block (huffman.go:40-107) ---
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
            TableSource ts{tree + 1, tree_avail - 1 < hb ? tree_avail - 1 : hb, 2};
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
    if (rc != SZB_OK) {
        if (lane == 0) a.lit_status[b] = rc;
        return;
    }

    // --- this block's streams (literals.go:270-371) ---
    const uint8_t *payload = a.src + d.src_off;
    uint32_t skip = d.lit_hdr_bytes + (d.lit_type == 2 ? tree_bytes : 0);
    int32_t comp = (int32_t)d.lit_comp - (int32_t)(d.lit_type == 2 ? tree_bytes : 0);
    uint8_t *out = a.litbuf + d.lit_buf_off;
    const uint32_t regen = d.lit_regen;
    int my_rc = SZB_OK;
    if (d.lit_streams == 1) {
        if (comp < 0)
            rc = SZB_ERR_PANIC;
        else if (lane == 0)
            my_rc = huf_decode_stream(sm.t.huf, max_bits, payload + skip, (uint32_t)comp, out, regen);
    } else {
        comp -= 6;
        if (comp < 0) {
            rc = SZB_ERR_PANIC;  // literals.go:283 negative slice bound
        } else {
            const uint8_t *jt = payload + skip;  // literals.go:46-58 jump table, 3 x u16 LE
            const uint32_t s1 = jt[0] | (jt[1] << 8), s2 = jt[2] | (jt[3] << 8), s3 = jt[4] | (jt[5] << 8);
            const uint32_t normal = (regen + 3) / 4;  // literals.go:306-311
            const int32_t last = (int32_t)regen - 3 * (int32_t)normal;
            if (s1 + s2 + s3 > (uint32_t)comp)
                rc = SZB_ERR_CORRUPTED_JUMPTABLE;  // literals.go:54-56 (the reference drops this error and re-slices)
            else if (last < 0)
                rc = SZB_ERR_PANIC;  // literals.go:311 inverted slice
            else if (lane < 4) {
                const uint32_t s4 = (uint32_t)comp - (s1 + s2 + s3);
                const uint32_t start = lane == 0 ? 0 : (lane == 1 ? s1 : (lane == 2 ? s1 + s2 : s1 + s2 + s3));
                const uint32_t len = lane == 0 ? s1 : (lane == 1 ? s2 : (lane == 2 ? s3 : s4));
                const uint32_t expected = lane < 3 ? normal : (uint32_t)last;
                my_rc = huf_decode_stream(sm.t.huf, max_bits, jt + 6 + start, len, out + lane * normal, expected);
            }
        }
    }
    if (rc == SZB_OK) rc = warp_first_error(my_rc);  // streams are checked in order 1..4 by the reference
    if (lane == 0) a.lit_status[b] = rc;
}

// shared memory per warp, k_sequences
struct SeqSmem {
    uint32_t tll[1 << kMaxALLL];
    uint32_t tml[1 << kMaxALML];
    uint32_t tof[1 << kMaxALOF];
    uint32_t buf[3][32];
    uint8_t symk[1 << kMaxALLL];
    int16_t norm[kMaxFseSymbols];
    uint16_t next[kMaxFseSymbols];
};

// One warp per compressed block that has sequences.
__global__ void __launch_bounds__(kCtaThreads) k_sequences(DeviceBatch a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kWarpsPerCta + warp_in_cta;
    if (w >= a.n_seq) return;
    SeqSmem &sm = reinterpret_cast<SeqSmem *>(smem_raw)[warp_in_cta];
    const uint32_t b = a.seq_list[w];
    const szb_block_desc d = a.blocks[b];
    const uint8_t *tables = a.src + d.src_off + d.seq_off + d.seq_hdr_bytes;
    const uint32_t tables_avail = d.block_size - d.seq_off - d.seq_hdr_bytes;

    // --- DecodeTables: LL, OF, ML in that order (sequences.go:275-369) ---
    uint32_t cursor = 0;
    uint32_t al[3] = {0, 0, 0};
    int rc = SZB_OK;
#pragma unroll
    for (int i = 0; i < 3 && rc == SZB_OK; i++) {
        const int kind = i;  // KIND_LL, KIND_OF, KIND_ML
        uint32_t *table = kind == KIND_LL ? sm.tll : (kind == KIND_OF ? sm.tof : sm.tml);
        const uint32_t mode = field_mode(d.seq_modes, kind);
        TableSource ts;
        if (mode == 3) {  // Repeat: rebuild from the origin block's bytes (host chased the chain)
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

    // --- DecodeSequences (sequences.go:126-206) ---
    const uint8_t *stream = tables + cursor;
    const uint32_t stream_len = tables_avail - cursor;
    RevBits r;
    SeqStates st{0, 0, 0};
    if (lane == 0) {
        if (!rev_init(r, stream, (int32_t)stream_len) || !rev_skip_padding(r)) {
            rc = SZB_ERR_BAD_PADDING;  // sequences.go:141-143
        } else {
            rev_refill(r);
            st.ll = rev_read(r, al[KIND_LL]);  // InitState order LL, OF, ML (sequences.go:145-159)
            st.of = rev_read(r, al[KIND_OF]);
            st.ml = rev_read(r, al[KIND_ML]);
        }
    }
    rc = __shfl_sync(kFull, rc, 0);
    if (rc != SZB_OK) {
        if (lane == 0) a.seq_status[b] = rc;
        return;
    }
    const uint32_t nseq = d.nseq;
    uint32_t *gll = a.seq_ll + d.seq_buf_off, *gml = a.seq_ml + d.seq_buf_off, *gof = a.seq_of + d.seq_buf_off;
    uint64_t ml_sum = 0;
    for (uint32_t base = 0; base < nseq; base += 32) {
        const uint32_t cnt = nseq - base < 32 ? nseq - base : 32;
        // pull the part of the backward bitstream the next rounds will consume towards L1
        {
            const uint32_t nxt = __shfl_sync(kFull, r.next, 0);
            const int64_t at = (int64_t)nxt - 128 * (int64_t)(lane + 1);
            if (lane < 4 && at >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(stream + at));
        }
        if (lane == 0) {
            for (uint32_t i = 0; i < cnt; i++) {
                uint32_t ll, ml, of;
                decode_one_sequence(r, sm.tll, sm.tof, sm.tml, st, base + i + 1 < nseq, &ll, &ml, &of);
                sm.buf[0][i] = ll;
                sm.buf[1][i] = ml;
                sm.buf[2][i] = of;
            }
        }
        __syncwarp();
        if (lane < cnt) {
            gll[base + lane] = sm.buf[0][lane];
            const uint32_t ml = sm.buf[1][lane];
            gml[base + lane] = ml;
            gof[base + lane] = sm.buf[2][lane];
            ml_sum += ml;
        }
        __syncwarp();
    }
    for (int dlt = 16; dlt > 0; dlt >>= 1) ml_sum += __shfl_xor_sync(kFull, ml_sum, dlt);
    if (lane == 0) {
        // the stream must be consumed exactly (sequences.go:197-204)
        if (r.remaining != 0) rc = SZB_ERR_NOT_ALL_BITS_USED;
        a.seq_status[b] = rc;
        a.out_size[b] = (uint64_t)d.lit_regen + ml_sum;
    }
}

// ---------------------------------------------------------------------------------------------
// Exclusive prefix sum of out_size over all blocks, one CTA.
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) k_scan_blocks(DeviceBatch a) {
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < a.nblocks; base += kScanThreads) {
        const uint32_t i = base + tid;
        const uint64_t v = i < a.nblocks ? a.out_size[i] : 0;
        uint64_t incl = v;
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            uint64_t t = __shfl_up_sync(kFull, incl, dlt);
            if ((int)lane >= dlt) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint64_t ws = warp_sums[lane];
            uint64_t wincl = ws;
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                uint64_t t = __shfl_up_sync(kFull, wincl, dlt);
                if ((int)lane >= dlt) wincl += t;
            }
            warp_sums[lane] = wincl - ws;  // exclusive over warps
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        const uint64_t excl = carry + warp_sums[wid] + incl - v;
        if (i < a.nblocks) a.out_off[i] = excl;
        __syncthreads();
        if (tid == kScanThreads - 1) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) a.total[0] = carry_s;
}

// ---------------------------------------------------------------------------------------------
// repeat-offset history (sequence_execution.go:65-114, table in SURVEY.md A.9)
struct History {
    uint32_t h0, h1, h2;
};
__device__ __forceinline__ uint32_t next_offset(History &h, uint32_t ofv, bool ll_zero) {
    uint32_t off;
    if (ofv > 3) {
        off = ofv - 3;
        h.h2 = h.h1;
        h.h1 = h.h0;
        h.h0 = off;
        return off;
    }
    const uint32_t idx = ofv - 1 + (ll_zero ? 1 : 0);  // 0: h0, 1: h1, 2: h2, 3: h0-1
    if (idx == 0) return h.h0;
    if (idx == 1) {
        off = h.h1;
        h.h1 = h.h0;
        h.h0 = off;
        return off;
    }
    off = idx == 2 ? h.h2 : h.h0 - 1;
    h.h2 = h.h1;
    h.h1 = h.h0;
    h.h0 = off;
    return off;
}

__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, uint64_t n, uint32_t lane) {
    for (uint64_t i = lane; i < n; i += 32) dst[i] = src[i];
}
__device__ __forceinline__ void warp_fill(uint8_t *dst, uint8_t v, uint64_t n, uint32_t lane) {
    for (uint64_t i = lane; i < n; i += 32) dst[i] = v;
}

constexpr uint32_t kLongLit = 32;
constexpr uint32_t kLongMatch = 64;

// One warp per frame; blocks in order; 32 sequences per round.
__global__ void __launch_bounds__(kCtaThreads) k_execute(DeviceBatch a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t f = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (f >= a.nframes) return;
    const szb_frame_desc fr = a.frames[f];
    const uint32_t b0 = fr.first_block, nb = fr.nblocks;
    uint8_t *const dst = a.dst;

    // the first failing block decides the frame's status, as in the sequential reference
    int err = SZB_OK;
    {
        uint32_t first_bad = 0xFFFFFFFFu;
        for (uint32_t i = lane; i < nb; i += 32) {
            if ((a.lit_status[b0 + i] | a.seq_status[b0 + i]) != 0 && i < first_bad) first_bad = i;
        }
        for (int dlt = 16; dlt > 0; dlt >>= 1) {
            uint32_t t = __shfl_xor_sync(kFull, first_bad, dlt);
            first_bad = t < first_bad ? t : first_bad;
        }
        if (first_bad != 0xFFFFFFFFu) {
            const int ls = a.lit_status[b0 + first_bad];
            err = ls ? ls : a.seq_status[b0 + first_bad];
        } else {
            err = fr.status;  // the header walk's verdict (blocks after the failing header are absent)
        }
    }
    const uint64_t frame_base = nb ? a.out_off[b0] : 0;
    uint64_t frame_len = 0;
    if (nb) frame_len = a.out_off[b0 + nb - 1] + a.out_size[b0 + nb - 1] - frame_base;
    if (err == SZB_OK && a.total[0] > a.dst_cap) err = SZB_ERR_DST_TOO_SMALL;
    if (err != SZB_OK) {
        if (lane == 0) {
            a.frame_status[f] = err;
            a.frame_out_off[f] = frame_base;
            a.frame_out_len[f] = 0;
        }
        return;
    }

    History hist{1, 4, 8};  // framedecompressor.go:48,59
    for (uint32_t bi = 0; bi < nb && err == SZB_OK; bi++) {
        const uint32_t b = b0 + bi;
        const szb_block_desc d = a.blocks[b];
        const uint8_t *payload = a.src + d.src_off;
        uint64_t out_pos = a.out_off[b];
        if (d.type == 0) {  // Raw (framedecompressor.go:211-215)
            warp_copy(dst + out_pos, payload, d.block_size, lane);
            __syncwarp();
            continue;
        }
        if (d.type == 1) {  // RLE (framedecompressor.go:229-241)
            warp_fill(dst + out_pos, payload[0], d.block_size, lane);
            __syncwarp();
            continue;
        }
        // Compressed: ExecuteSequences (sequence_execution.go:14-63)
        const bool lit_rle = d.lit_type == 1;
        const uint8_t *lit = d.lit_type == 0 ? payload + d.lit_hdr_bytes : a.litbuf + d.lit_buf_off;
        const uint8_t rle_byte = lit_rle ? payload[d.lit_hdr_bytes] : 0;
        const uint32_t nseq = d.nseq;
        const uint32_t *gll = a.seq_ll + d.seq_buf_off, *gml = a.seq_ml + d.seq_buf_off, *gof = a.seq_of + d.seq_buf_off;
        uint32_t lit_pos = 0;
        for (uint32_t base = 0; base < nseq; base += 32) {
            const uint32_t cnt = nseq - base < 32 ? nseq - base : 32;
            const bool act = lane < cnt;
            const uint32_t ll = act ? gll[base + lane] : 0;
            const uint32_t ml = act ? gml[base + lane] : 0;
            const uint32_t ofv = act ? gof[base + lane] : 4;

            // --- offsets through the 3-entry history (nextOffset) ---
            uint32_t off;
            const uint32_t rep_mask = __ballot_sync(kFull, act && ofv <= 3);
            if (rep_mask == 0) {
                off = ofv - 3;
                const uint32_t o1 = __shfl_sync(kFull, off, cnt - 1);
                const uint32_t o2 = __shfl_sync(kFull, off, cnt >= 2 ? cnt - 2 : 0);
                const uint32_t o3 = __shfl_sync(kFull, off, cnt >= 3 ? cnt - 3 : 0);
                if (cnt >= 3) {
                    hist = History{o1, o2, o3};
                } else if (cnt == 2) {
                    hist = History{o1, o2, hist.h0};
                } else {
                    hist = History{o1, hist.h0, hist.h1};
                }
            } else {
                off = 0;
                for (uint32_t j = 0; j < cnt; j++) {  // serial, every lane tracks the same history
                    const uint32_t v = __shfl_sync(kFull, ofv, j);
                    const uint32_t l = __shfl_sync(kFull, ll, j);
                    const uint32_t o = next_offset(hist, v, l == 0);
                    if (lane == j) off = o;
                }
            }

            // --- positions: prefix sums over the round ---
            const uint32_t tot = ll + ml;
            uint32_t incl_ll = ll, incl_tot = tot;
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                const uint32_t t1 = __shfl_up_sync(kFull, incl_ll, dlt);
                const uint32_t t2 = __shfl_up_sync(kFull, incl_tot, dlt);
                if ((int)lane >= dlt) {
                    incl_ll += t1;
                    incl_tot += t2;
                }
            }
            const uint32_t round_ll = __shfl_sync(kFull, incl_ll, 31);
            const uint32_t round_tot = __shfl_sync(kFull, incl_tot, 31);
            if ((uint64_t)lit_pos + round_ll > d.lit_regen) {
                // literals.go:398-409 Read runs dry / sequence_execution.go:26-28; RLE literals: GetRest panics
                err = lit_rle ? SZB_ERR_PANIC : SZB_ERR_DIDNT_COPY_ALL_LITERAL_BYTES;
                break;
            }
            const uint64_t my_dst = out_pos + (incl_tot - tot);
            const uint32_t my_lit = lit_pos + (incl_ll - ll);

            // --- literal runs (sequence_execution.go:19-34) ---
            if (act && ll > 0 && ll < kLongLit) {
                if (lit_rle) {
                    for (uint32_t k = 0; k < ll; k++) dst[my_dst + k] = rle_byte;
                } else {
                    for (uint32_t k = 0; k < ll; k++) dst[my_dst + k] = lit[my_lit + k];
                }
            }
            uint32_t long_lit = __ballot_sync(kFull, act && ll >= kLongLit);
            while (long_lit) {
                const int j = __ffs(long_lit) - 1;
                long_lit &= long_lit - 1;
                const uint32_t L = __shfl_sync(kFull, ll, j);
                const uint64_t D = __shfl_sync(kFull, my_dst, j);
                const uint32_t S = __shfl_sync(kFull, my_lit, j);
                if (lit_rle)
                    warp_fill(dst + D, rle_byte, L, lane);
                else
                    warp_copy(dst + D, lit + S, L, lane);
            }

            // --- matches (RepeatBeforeIndex, ringbuffer.go:242-277) in dependency rounds ---
            const uint64_t mdst = my_dst + ll;
            const bool has_match = act && ml > 0;
            if (__any_sync(kFull, has_match && (off == 0 || (uint64_t)off > mdst - frame_base))) {
                err = SZB_ERR_CANT_REPEAT_BYTES;  // ringbuffer.go:203-214
                break;
            }
            const uint64_t msrc = mdst - off;
            const uint64_t need_end = (msrc + ml < mdst) ? msrc + ml : mdst;  // bytes other lanes may still owe us
            __syncwarp();
            uint32_t pending = __ballot_sync(kFull, has_match);
            while (pending) {
                const int first = __ffs(pending) - 1;
                const uint64_t frontier = __shfl_sync(kFull, mdst, first);  // everything below is final
                const uint32_t first_ml = __shfl_sync(kFull, ml, first);
                if (first_ml >= kLongMatch) {  // warp-wide copy of one long match
                    const uint32_t OFF = __shfl_sync(kFull, off, first);
                    const uint8_t *S = dst + frontier - OFF;
                    uint8_t *D = dst + frontier;
                    if (OFF >= 32) {
                        for (uint32_t k0 = 0; k0 < first_ml; k0 += 32) {
                            const uint32_t k = k0 + lane;
                            if (k < first_ml) D[k] = S[k];
                            __syncwarp();
                        }
                    } else {  // overlapping: periodic extension of the OFF bytes before the match
                        for (uint32_t k = lane; k < first_ml; k += 32) D[k] = S[k % OFF];
                    }
                    pending &= ~(1u << first);
                    __syncwarp();
                    continue;
                }
                const bool ready = ((pending >> lane) & 1) && ml < kLongMatch && ((int)lane == first || need_end <= frontier);
                if (ready) {
                    for (uint32_t k = 0; k < ml; k++) dst[mdst + k] = dst[msrc + k];  // byte-serial: handles self overlap
                }
                pending &= ~__ballot_sync(kFull, ready);
                __syncwarp();
            }
            out_pos += round_tot;
            lit_pos += round_ll;
        }
        if (err != SZB_OK) break;
        // trailing literals (sequence_execution.go:55-60, literals.go:411-420)
        const uint32_t rest = d.lit_regen - lit_pos;
        if (lit_rle)
            warp_fill(dst + out_pos, rle_byte, rest, lane);
        else
            warp_copy(dst + out_pos, lit + lit_pos, rest, lane);
        __syncwarp();
    }
    if (lane == 0) {
        a.frame_status[f] = err;
        a.frame_out_off[f] = frame_base;
        a.frame_out_len[f] = err == SZB_OK ? frame_len : 0;
    }
}

}  // namespace szb
