// execute_long.cuh -- stage 4 for LONG frames, parallel over the blocks and bytes of a frame (sm_100a).
//
// The reference executes a frame strictly in order: sequence after sequence through a ring buffer
// (decompression/sequence_execution.go:14-63, ringbuffer.go:197-277), the repeat-offset history carried from block to
// block (sequence_execution.go:65-114).  k_execute keeps that order inside a frame (one warp per frame) and is parallel
// over frames; a frame of gigabytes then runs on one warp.  The kernels here remove the two things that are sequential:
//
//   k_long_hist     the HISTORY: every block (or SLICE of a block: DeviceBatch::long_slice sequences, one warp each) is walked
//                   with a symbolic history (an entry is a constant, or "entry i of the history the slice started with, minus
//                   k") -- its transfer function; the sums of its literal and match lengths come out of the same walk;
//   k_long_blockscan  one warp per block scans the functions and sums of the block's slices: what lies between the block's
//                   start and every slice, and the block's whole function;
//   k_long_compose  one warp per frame composes the blocks' functions in block order: the history every block starts with;
//   k_long_emit     one warp per slice, every slice at once: offsets through the (now known) history, positions by
//                   prefix sums; writes the literal bytes to their place and, for EVERY output byte, a DISTANCE cell:
//                   0 for a literal byte, else how far back the byte's source lies (inside an overlapping match a
//                   multiple of the offset, so that the source lies in front of the match: ringbuffer.go:236-262);
//   k_long_jump     the COPY ORDER: one thread per output byte follows d[p] += d[p - d[p]] until p - d[p] is a literal
//                   byte and copies it.  Cells are updated in place after every step; whatever value a cell holds at
//                   any time points at an ancestor of its byte, so no ordering between bytes, no flags and no waiting
//                   are needed -- every byte in flight shortens the walk of the bytes that hang on it (pointer doubling),
//                   and bytes handled earlier have their literal one step away (tiles run in address order);
//   k_long_verdict  errors found while emitting (bad offset, literals run dry) become the frame's status: the first
//                   failing block and round decides, as in the sequential reference.
//
// Frames this path cannot take (more than 4 GiB of output, more output than the host's bound, no scratch memory)
// stay on k_execute_pair; every kernel evaluates the same predicate (long_jump_ok).
#pragma once
// Included by kernels.cuh, inside namespace szb, after DeviceBatch and the stage-4 helpers.

#ifndef SZB_EMIT_LD
#define SZB_EMIT_LD 1  // 1 = k_long_emit reads back its own cells through L1 (same warp, after __syncwarp): 2 % faster than ld.cg
#endif
#ifndef SZB_JUMP_LD
#define SZB_JUMP_LD 1  // 1 = k_long_jump follows cells through L1 (a stale cell is still an ancestor): 9.6 -> 6.9 ms on one 256 MiB frame
#endif
#ifndef SZB_EMIT_CTAS
#define SZB_EMIT_CTAS 8  // CTAs per SM k_long_emit is compiled for (64 registers)
#endif
#ifndef SZB_JUMP_BATCHED
#define SZB_JUMP_BATCHED 1  // 1 = the loads of a thread's walks are issued together (predicated), then consumed; 0 = a branch per walk
#endif
#ifndef SZB_JUMP_REFILL
#define SZB_JUMP_REFILL 0  // experiment, see k_long_jump
#endif
#ifndef SZB_JUMP_CTAS_PER_SM
#define SZB_JUMP_CTAS_PER_SM 5  // 8 walks per thread need ~48 registers: 5 CTAs of 256 threads, 10 240 loads in flight per SM
#endif
__device__ __forceinline__ uint32_t ld_ca(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
    uint32_t v;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ uint32_t emit_ld(const uint32_t *p) {
#if SZB_EMIT_LD
    return ld_ca(p);
#else
    return __ldcg(p);
#endif
}
// Predicated loads: the load is issued only where `on`, without a branch, so that the loads of a thread's independent
// walks leave back to back (a branch per walk makes the compiler wait for each load before it issues the next).
__device__ __forceinline__ uint32_t jump_ld_if(const uint32_t *p, uint32_t on) {
#if defined(__CUDA_ARCH__)
    uint32_t v;
#if SZB_JUMP_LD
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.global.ca.u32 %0, [%1];\n\t}" : "=r"(v) : "l"(p), "r"(on) : "memory");
#else
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.global.cg.u32 %0, [%1];\n\t}" : "=r"(v) : "l"(p), "r"(on) : "memory");
#endif
    return v;
#else
    return on ? *p : 0;
#endif
}
__device__ __forceinline__ uint32_t ld_u8_if(const uint8_t *p, uint32_t on) {
#if defined(__CUDA_ARCH__)
    uint32_t v;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.global.u8 %0, [%1];\n\t}" : "=r"(v) : "l"(p), "r"(on) : "memory");
    return v;
#else
    return on ? *p : 0;
#endif
}
__device__ __forceinline__ uint32_t jump_ld(const uint32_t *p) {
#if SZB_JUMP_LD
    return ld_ca(p);
#else
    return __ldcg(p);
#endif
}
#ifndef SZB_JUMP_CHAINS
#define SZB_JUMP_CHAINS 8  // 8: 3-4 % faster than 4 (profiles/README.md, r01k)
#endif
constexpr int kJumpThreads = 256;
constexpr int kJumpChains = SZB_JUMP_CHAINS;  // bytes (independent walks in flight) per thread of k_long_jump
#ifndef SZB_JUMP_TILE
#define SZB_JUMP_TILE 512  // 512: stage 4 of one 256 MiB frame 6.9 -> 6.0 ms against 1024 (fewer claimed-but-unstarted cells); 256: the same
#endif
constexpr uint32_t kJumpTile = SZB_JUMP_TILE;  // cells per ticket of k_long_jump; every frame's cells start at a multiple of it
static_assert(kJumpTile % (32 * kJumpChains) == 0, "a tile is a whole number of warp steps");
constexpr unsigned long long kLongNoError = ~0ull;

__device__ __forceinline__ bool long_jump_ok(const DeviceBatch &a, uint32_t slot) {
    if (a.dist == nullptr) return false;
    const uint32_t f = a.exec_list[slot];
    if (a.frame_status[f] != SZB_OK) return false;
    const uint64_t len = a.frame_out_len[f];
    return len <= a.long_dbase[slot + 1] - a.long_dbase[slot] && len <= 0x100000000ull;  // a distance is < len: 32 bits
}

// ---- history values: concrete (uint32_t) or symbolic (uint64_t) ----
constexpr uint64_t kSymBit = 1ull << 63;
__device__ __forceinline__ uint64_t sym_entry(uint32_t idx) { return kSymBit | ((uint64_t)idx << 32); }
__device__ __forceinline__ uint32_t hist_dec1(uint32_t v) { return v - 1; }
__device__ __forceinline__ uint64_t hist_dec1(uint64_t v) { return (v & kSymBit) ? v + 1 : (uint64_t)((uint32_t)v - 1u); }
__device__ __forceinline__ uint32_t hist_apply(uint64_t t, uint32_t h0, uint32_t h1, uint32_t h2) {
    if (!(t & kSymBit)) return (uint32_t)t;
    const uint32_t idx = (uint32_t)(t >> 32) & 3;
    return (idx == 0 ? h0 : (idx == 1 ? h1 : h2)) - (uint32_t)t;
}

template <class V>
struct HistV {
    V h0, h1, h2;
};

// sequence_execution.go:65-114 (the table in SURVEY.md A.9) for a repeat code (ofv 1..3)
template <class V>
__device__ __forceinline__ V next_repeat(HistV<V> &h, uint32_t ofv, bool ll_zero) {
    const uint32_t idx = ofv - 1 + (ll_zero ? 1 : 0);  // 0: h0, 1: h1, 2: h2, 3: h0 - 1
    if (idx == 0) return h.h0;
    V off;
    if (idx == 1) {
        off = h.h1;
        h.h1 = h.h0;
        h.h0 = off;
        return off;
    }
    off = idx == 2 ? h.h2 : hist_dec1(h.h0);
    h.h2 = h.h1;
    h.h1 = h.h0;
    h.h0 = off;
    return off;
}

// One round of up to 32 sequences (lane = sequence) through the history.  A sequence with a direct offset pushes it
// whatever the history holds, so the walk goes from repeat code to repeat code and folds the direct runs in between.
// Returns the lane's offset (of its own sequence).
template <class V>
__device__ __forceinline__ V walk_round(HistV<V> &hist, uint32_t ofv, uint32_t ll, bool act, uint32_t cnt, uint32_t lane) {
    V off = (V)(ofv - 3);
    uint32_t rm = __ballot_sync(kFull, act && ofv <= 3);
    uint32_t p = 0;  // sequences [0, p) are folded into hist
    for (;;) {
        const uint32_t j = rm ? (uint32_t)__ffs(rm) - 1 : cnt;  // the next repeat code, or the end of the round
        const uint32_t n = j - p;
        if (n) {
            const V o1 = (V)(__shfl_sync(kFull, ofv, j - 1) - 3);
            const V o2 = (V)(__shfl_sync(kFull, ofv, n >= 2 ? j - 2 : 0) - 3);
            const V o3 = (V)(__shfl_sync(kFull, ofv, n >= 3 ? j - 3 : 0) - 3);
            if (n >= 3) {
                hist = HistV<V>{o1, o2, o3};
            } else if (n == 2) {
                hist = HistV<V>{o1, o2, hist.h0};
            } else {
                hist = HistV<V>{o1, hist.h0, hist.h1};
            }
        }
        if (j >= cnt) break;
        const uint32_t v = __shfl_sync(kFull, ofv, j);
        const uint32_t l = __shfl_sync(kFull, ll, j);
        const V o = next_repeat(hist, v, l == 0);  // every lane tracks the same history
        if (lane == j) off = o;
        rm &= rm - 1;
        p = j + 1;
    }
    return off;
}

// ---- k_long_hist: the transfer function and the length sums of every SLICE of the long frames' blocks; one warp each ----
// A slice is a run of DeviceBatch::long_slice sequences of one block (the whole block when long_slice is 0).
__device__ __forceinline__ uint32_t slice_end(const DeviceBatch &a, uint32_t seq0, uint32_t nseq) {
    return a.long_slice && nseq - seq0 > a.long_slice ? seq0 + a.long_slice : nseq;
}
__global__ void __launch_bounds__(kCtaThreads) k_long_hist(DeviceBatch a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (w >= a.n_ls) return;
    const uint32_t b = a.lb_block[a.ls_lb[w]];
    const szb_block_desc d = a.blocks[b];
    HistV<uint64_t> h{sym_entry(0), sym_entry(1), sym_entry(2)};
    uint64_t sum_ll = 0, sum_tot = 0;  // this lane's share
    if (d.type == 2 && d.nseq > 0 && a.frame_status[d.frame] == SZB_OK) {
        const uint32_t seq0 = a.ls_seq0[w], seq1 = slice_end(a, seq0, d.nseq);
        const uint32_t *tr = a.seq_ll + (d.seq_buf_off + lane);  // the arrays are padded to whole rounds
        uint32_t ll = tr[seq0], ml = tr[seq0 + a.seq_stride], ofv = tr[seq0 + 2 * a.seq_stride];
        for (uint32_t base = seq0; base < seq1; base += 32) {
            const uint32_t cnt = seq1 - base < 32 ? seq1 - base : 32;
            const bool act = lane < cnt;
            const uint32_t c_ll = ll, c_ofv = act ? ofv : 4;
            if (act) {
                sum_ll += ll;
                sum_tot += (uint64_t)ll + ml;
            }
            if (base + 32 < seq1) {  // the next round's loads are in flight during the walk
                ll = tr[base + 32];
                ml = tr[base + 32 + a.seq_stride];
                ofv = tr[base + 32 + 2 * a.seq_stride];
            }
            walk_round<uint64_t>(h, c_ofv, c_ll, act, cnt, lane);
        }
    }
    for (int dlt = 16; dlt; dlt >>= 1) {
        sum_ll += __shfl_xor_sync(kFull, sum_ll, dlt);
        sum_tot += __shfl_xor_sync(kFull, sum_tot, dlt);
    }
    if (lane == 0) {
        a.ls_T[3 * (uint64_t)w] = h.h0;
        a.ls_T[3 * (uint64_t)w + 1] = h.h1;
        a.ls_T[3 * (uint64_t)w + 2] = h.h2;
        a.ls_sum[2 * (uint64_t)w] = sum_ll;
        a.ls_sum[2 * (uint64_t)w + 1] = sum_tot;
    }
}

// an entry of a LATER transfer function over the entries (a0, a1, a2) of the function before it
__device__ __forceinline__ uint64_t sym_compose(uint64_t e, uint64_t a0, uint64_t a1, uint64_t a2) {
    if (!(e & kSymBit)) return e;
    const uint32_t idx = (uint32_t)(e >> 32) & 3;
    const uint64_t base = idx == 0 ? a0 : (idx == 1 ? a1 : a2);
    if (base & kSymBit) return base + (uint32_t)e;  // "entry j minus m" minus k
    return (uint64_t)((uint32_t)base - (uint32_t)e);
}

// ---- k_long_blockscan: one warp per block: the slices' functions and sums scanned in slice order ----
// Afterwards a slice's entry holds what lies between the block's start and its own: the composed transfer function and the
// literal / output bytes in front of it; long_T gets the block's whole function (what k_long_compose scans over the frame).
__global__ void __launch_bounds__(kCtaThreads) k_long_blockscan(DeviceBatch a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lb = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (lb >= a.n_lb) return;
    const uint32_t first = a.lb_first_ls[lb], end = a.lb_first_ls[lb + 1];
    uint64_t c0 = sym_entry(0), c1 = sym_entry(1), c2 = sym_entry(2);  // everything before this step, as a function of the block's start
    uint64_t cs_ll = 0, cs_tot = 0;
    for (uint32_t base = first; base < end; base += 32) {
        const uint32_t i = base + lane;
        uint64_t t0 = sym_entry(0), t1 = sym_entry(1), t2 = sym_entry(2), s_ll = 0, s_tot = 0;
        if (i < end) {
            t0 = a.ls_T[3 * (uint64_t)i];
            t1 = a.ls_T[3 * (uint64_t)i + 1];
            t2 = a.ls_T[3 * (uint64_t)i + 2];
            s_ll = a.ls_sum[2 * (uint64_t)i];
            s_tot = a.ls_sum[2 * (uint64_t)i + 1];
        }
        for (uint32_t dlt = 1; dlt < 32; dlt <<= 1) {  // inclusive scans
            const uint64_t g0 = __shfl_up_sync(kFull, t0, dlt), g1 = __shfl_up_sync(kFull, t1, dlt), g2 = __shfl_up_sync(kFull, t2, dlt);
            const uint64_t u_ll = __shfl_up_sync(kFull, s_ll, dlt), u_tot = __shfl_up_sync(kFull, s_tot, dlt);
            if (lane >= dlt) {
                t0 = sym_compose(t0, g0, g1, g2);
                t1 = sym_compose(t1, g0, g1, g2);
                t2 = sym_compose(t2, g0, g1, g2);
                s_ll += u_ll;
                s_tot += u_tot;
            }
        }
        uint64_t e0 = __shfl_up_sync(kFull, t0, 1), e1 = __shfl_up_sync(kFull, t1, 1), e2 = __shfl_up_sync(kFull, t2, 1);
        uint64_t x_ll = __shfl_up_sync(kFull, s_ll, 1), x_tot = __shfl_up_sync(kFull, s_tot, 1);
        if (lane == 0) e0 = sym_entry(0), e1 = sym_entry(1), e2 = sym_entry(2), x_ll = 0, x_tot = 0;
        if (i < end) {
            a.ls_T[3 * (uint64_t)i] = sym_compose(e0, c0, c1, c2);
            a.ls_T[3 * (uint64_t)i + 1] = sym_compose(e1, c0, c1, c2);
            a.ls_T[3 * (uint64_t)i + 2] = sym_compose(e2, c0, c1, c2);
            a.ls_sum[2 * (uint64_t)i] = cs_ll + x_ll;
            a.ls_sum[2 * (uint64_t)i + 1] = cs_tot + x_tot;
        }
        const uint64_t l0 = __shfl_sync(kFull, t0, 31), l1 = __shfl_sync(kFull, t1, 31), l2 = __shfl_sync(kFull, t2, 31);
        const uint64_t n0 = sym_compose(l0, c0, c1, c2), n1 = sym_compose(l1, c0, c1, c2), n2 = sym_compose(l2, c0, c1, c2);
        c0 = n0;
        c1 = n1;
        c2 = n2;
        cs_ll += __shfl_sync(kFull, s_ll, 31);
        cs_tot += __shfl_sync(kFull, s_tot, 31);
    }
    if (lane == 0) {
        a.long_T[3 * (uint64_t)lb] = c0;
        a.long_T[3 * (uint64_t)lb + 1] = c1;
        a.long_T[3 * (uint64_t)lb + 2] = c2;
    }
}

// ---- k_long_compose: the history each block starts with; one warp per long frame ----
// Transfer functions compose associatively: 32 blocks per step by a warp scan, the history carried from step to step.
__global__ void __launch_bounds__(32) k_long_compose(DeviceBatch a) {
    const uint32_t lane = threadIdx.x;
    const uint32_t slot = blockIdx.x;
    if (slot >= a.n_long) return;
    const uint32_t first = a.long_first_lb[slot], end = a.long_first_lb[slot + 1];
    uint32_t h0 = 1, h1 = 4, h2 = 8;  // framedecompressor.go:48,59
    uint64_t n0 = sym_entry(0), n1 = sym_entry(1), n2 = sym_entry(2);
    if (first + lane < end) {
        n0 = a.long_T[3 * (uint64_t)(first + lane)];
        n1 = a.long_T[3 * (uint64_t)(first + lane) + 1];
        n2 = a.long_T[3 * (uint64_t)(first + lane) + 2];
    }
    for (uint32_t base = first; base < end; base += 32) {
        const uint32_t i = base + lane;
        uint64_t t0 = n0, t1 = n1, t2 = n2;  // blocks past the end hold the identity
        n0 = sym_entry(0), n1 = sym_entry(1), n2 = sym_entry(2);
        if (i + 32 < end) {  // the next step's functions are on their way during the scan
            n0 = a.long_T[3 * (uint64_t)(i + 32)];
            n1 = a.long_T[3 * (uint64_t)(i + 32) + 1];
            n2 = a.long_T[3 * (uint64_t)(i + 32) + 2];
        }
        // inclusive scan: lane j ends up with the function of blocks base .. base + j
        for (uint32_t dlt = 1; dlt < 32; dlt <<= 1) {
            const uint64_t g0 = __shfl_up_sync(kFull, t0, dlt), g1 = __shfl_up_sync(kFull, t1, dlt), g2 = __shfl_up_sync(kFull, t2, dlt);
            if (lane >= dlt) {
                t0 = sym_compose(t0, g0, g1, g2);
                t1 = sym_compose(t1, g0, g1, g2);
                t2 = sym_compose(t2, g0, g1, g2);
            }
        }
        // a block starts with what the blocks before it leave behind
        uint64_t e0 = __shfl_up_sync(kFull, t0, 1), e1 = __shfl_up_sync(kFull, t1, 1), e2 = __shfl_up_sync(kFull, t2, 1);
        if (lane == 0) e0 = sym_entry(0), e1 = sym_entry(1), e2 = sym_entry(2);
        if (i < end) {
            a.long_hist[3 * (uint64_t)i] = hist_apply(e0, h0, h1, h2);
            a.long_hist[3 * (uint64_t)i + 1] = hist_apply(e1, h0, h1, h2);
            a.long_hist[3 * (uint64_t)i + 2] = hist_apply(e2, h0, h1, h2);
        }
        const uint64_t l0 = __shfl_sync(kFull, t0, 31), l1 = __shfl_sync(kFull, t1, 31), l2 = __shfl_sync(kFull, t2, 31);
        const uint32_t c0 = hist_apply(l0, h0, h1, h2), c1 = hist_apply(l1, h0, h1, h2), c2 = hist_apply(l2, h0, h1, h2);
        h0 = c0;
        h1 = c1;
        h2 = c2;
    }
}

// ---- k_long_emit: literal bytes and distance cells of every slice of the long frames' blocks; one warp per slice ----
__global__ void __launch_bounds__(kCtaThreads, SZB_EMIT_CTAS) k_long_emit(DeviceBatch a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (w >= a.n_ls) return;
    const uint32_t lb = a.ls_lb[w];
    const uint32_t slot = a.lb_slot[lb];
    if (!long_jump_ok(a, slot)) return;
    const uint32_t b = a.lb_block[lb];
    const szb_block_desc d = a.blocks[b];
    const uint64_t frame_base = a.frame_out_off[d.frame];
    uint8_t *const dst = a.dst;
    uint32_t *const cells = a.dist + a.long_dbase[slot];  // cell of the frame's first byte

    if (d.type != 2 || d.nseq == 0) {
        // Raw / RLE bodies and blocks without sequences: written by k_execute_bodies; every byte is a literal byte
        const uint64_t n = a.out_size[b], at = a.out_off[b] - frame_base;
        for (uint64_t i = lane; i < n; i += 32) cells[at + i] = 0;
        return;
    }
    const uint8_t *payload = a.src + d.src_off;
    const bool lit_rle = d.lit_type == 1;
    const uint32_t fill = lit_rle ? payload[d.lit_hdr_bytes] : 0;
    const uint8_t *__restrict__ lit = d.lit_type == 0 ? payload + d.lit_hdr_bytes : a.litbuf + d.lit_buf_off;  // unused for RLE literals
    const uint32_t nseq = d.nseq;
    const uint32_t seq0 = a.ls_seq0[w], seq1 = slice_end(a, seq0, nseq);
    // what the slices of the block in front of this one leave behind (k_long_blockscan), from what the block starts with (k_long_compose)
    const uint32_t b0 = a.long_hist[3 * (uint64_t)lb], b1 = a.long_hist[3 * (uint64_t)lb + 1], b2 = a.long_hist[3 * (uint64_t)lb + 2];
    HistV<uint32_t> hist{hist_apply(a.ls_T[3 * (uint64_t)w], b0, b1, b2), hist_apply(a.ls_T[3 * (uint64_t)w + 1], b0, b1, b2),
                         hist_apply(a.ls_T[3 * (uint64_t)w + 2], b0, b1, b2)};
    const uint64_t lit_before = a.ls_sum[2 * (uint64_t)w], out_before = a.ls_sum[2 * (uint64_t)w + 1];
    int err = SZB_OK;
    uint32_t err_base = 0;
    if (lit_before > d.lit_regen) {
        // an earlier slice of the block runs out of literals (and says so with a smaller key): nothing of this one may be emitted
        if (lane == 0)
            atomicMin(&a.long_err[slot], ((unsigned long long)lb << 40) | ((unsigned long long)seq0 << 8) |
                                             (uint32_t)(-(lit_rle ? SZB_ERR_PANIC : SZB_ERR_DIDNT_COPY_ALL_LITERAL_BYTES)));
        return;
    }
    uint32_t lit_pos = (uint32_t)lit_before;
    uint64_t out_pos = a.out_off[b] + out_before;
    const uint64_t fb0 = out_pos - frame_base;  // frame bytes in front of the slice
    uint64_t fb = fb0;                          // frame bytes in front of the round
    const uint32_t *tr = a.seq_ll + (d.seq_buf_off + lane);
    uint32_t n_ll = tr[seq0], n_ml = tr[seq0 + a.seq_stride], n_ofv = tr[seq0 + 2 * a.seq_stride];
    for (uint32_t base = seq0; base < seq1; base += 32) {
        const uint32_t cnt = seq1 - base < 32 ? seq1 - base : 32;
        const bool act = lane < cnt;
        const uint32_t ll = act ? n_ll : 0, ml = act ? n_ml : 0, ofv = act ? n_ofv : 4;
        if (base + 32 < seq1) {
            n_ll = tr[base + 32];
            n_ml = tr[base + 32 + a.seq_stride];
            n_ofv = tr[base + 32 + 2 * a.seq_stride];
        }
        const uint32_t off = walk_round<uint32_t>(hist, ofv, ll, act, cnt, lane);
        // positions: prefix sums over the round (a sequence is < 2^18 bytes: the sums fit 32 bits)
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
            err_base = base;
            break;
        }
        const uint32_t excl_tot = incl_tot - tot, excl_ll = incl_ll - ll;
        if (!lit_rle && lane < 4) {  // the literals of the next rounds are on their way to L1 while this round is emitted
            const uint32_t ahead = lit_pos + round_ll + 128 * lane;
            if (ahead < d.lit_regen) SZB_PREFETCH_L1(lit + ahead);
        }
        // every match must lie inside the frame (ringbuffer.go:203-214); a match length of 0 cannot come out of stage 3
        if (__any_sync(kFull, act && ((uint64_t)off > fb + excl_tot + ll || off == 0 || ml == 0))) {
            err = SZB_ERR_CANT_REPEAT_BYTES;
            err_base = base;
            break;
        }
        // Every byte of the round, 32 consecutive bytes per step in address order: its sequence by binary search over the
        // inclusive sums (lanes past cnt repeat the last sum).  A source inside the block that an EARLIER step (or round)
        // of this warp emitted has its cell in memory already, and that cell is resolved as far as the block allows:
        // adding it makes this cell point at a literal byte of the block or in front of the block, so that the chains
        // k_long_jump still has to follow go from block to block instead of from match to match.
        for (uint32_t x0 = 0; x0 < round_tot; x0 += 32) {
            const uint32_t x = x0 + lane;
            uint32_t j = 0;
#pragma unroll
            for (uint32_t step = 16; step; step >>= 1) {
                const uint32_t t = __shfl_sync(kFull, incl_tot, j + step - 1);
                if (t <= x) j += step;
            }
            const uint32_t s_excl = __shfl_sync(kFull, excl_tot, j);
            const uint32_t s_ll = __shfl_sync(kFull, ll, j);
            const uint32_t s_off = __shfl_sync(kFull, off, j);
            const uint32_t s_lit = __shfl_sync(kFull, excl_ll, j);
            __syncwarp();  // the cells of the steps before this one are in memory
            if (x < round_tot) {
                const uint32_t k = x - s_excl;
                uint32_t cell = 0;
                if (k < s_ll) {
                    dst[out_pos + x] = lit_rle ? (uint8_t)fill : lit[lit_pos + s_lit + k];
                } else {
                    const uint32_t m = k - s_ll;
                    cell = m < s_off ? s_off : s_off * (m / s_off + 1);
                    const uint64_t me = fb + x;  // my byte, counted from the frame's start
                    if (cell <= me - fb0 && cell > lane) cell += emit_ld(cells + (me - cell));  // in the slice, below this step
                }
                cells[fb + x] = cell;
            }
        }
        out_pos += round_tot;
        fb += round_tot;
        lit_pos += round_ll;
    }
    if (err != SZB_OK) {
        if (lane == 0) atomicMin(&a.long_err[slot], ((unsigned long long)lb << 40) | ((unsigned long long)err_base << 8) | (uint32_t)(-err));
        return;
    }
    if (seq1 != nseq) return;
    // trailing literals (sequence_execution.go:55-60, literals.go:411-420): after the block's last sequence
    const uint32_t rest = d.lit_regen - lit_pos;
    for (uint32_t i = lane; i < rest; i += 32) {
        dst[out_pos + i] = lit_rle ? (uint8_t)fill : lit[lit_pos + i];
        cells[fb + i] = 0;
    }
}

// ---- k_long_jump: every match byte finds its literal byte ----
// Tiles of kJumpTile cells are handed out in address order from a ticket counter, one tile per warp at a time: whatever
// the number of resident warps (or the kernels running beside this one), the tiles in flight are the lowest unfinished
// ones, so that a walk leaves the region in flight after a few steps and lands on a finished cell.  (A static
// tile-to-CTA map loses that as soon as one CTA of the grid is not resident: 6.8 -> 10.2 ms on one 256 MiB frame.)
__global__ void __launch_bounds__(kJumpThreads, SZB_JUMP_CTAS_PER_SM) k_long_jump(DeviceBatch a) {
    const uint64_t total = a.long_dbase[a.n_long];
    const uint32_t lane = threadIdx.x & 31;
    unsigned long long next = 0;
    if (lane == 0) next = atomicAdd(a.long_ticket, 1ull);
    for (;;) {
        const uint64_t c0 = __shfl_sync(kFull, next, 0) * kJumpTile;
        if (c0 >= total) break;
        if (lane == 0) next = atomicAdd(a.long_ticket, 1ull);  // the next tile's ticket is on its way meanwhile
        // the frame of the tile: the last slot whose cells start at or before c0
        uint32_t lo = 0, hi = a.n_long;  // invariant: dbase[lo] <= c0 < dbase[hi]
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (a.long_dbase[mid] <= c0)
                lo = mid;
            else
                hi = mid;
        }
        const uint32_t slot = lo;
        if (!long_jump_ok(a, slot) || a.long_err[slot] != kLongNoError) continue;
        const uint32_t f = a.exec_list[slot];
        const uint64_t len = a.frame_out_len[f];
        const uint64_t rel0 = c0 - a.long_dbase[slot];
        if (rel0 >= len) continue;
        uint32_t *const cells = a.dist + a.long_dbase[slot];
        uint8_t *const out = a.dst + a.frame_out_off[f];
        // positions inside the frame fit 32 bits (long_jump_ok), and so does the whole tile: rel0 is a multiple of the
        // tile and below len <= 2^32
#if SZB_JUMP_REFILL
        // EXPERIMENT (not measured yet): a lane owns the bytes lane, lane + 32, ... of the tile and keeps kJumpChains walks
        // going; a walk that ends hands its slot to the lane's next byte at once, instead of the lane idling until the
        // slowest walk of the warp's step is over (with one step per 32 x kJumpChains bytes, 14 of 32 lanes are busy on
        // average).  A walk starts at distance 0, so that its first load is the byte's own cell; the bytes are copied in a
        // second, fully batched pass over the (now final) cells.
        {
            const uint32_t base = (uint32_t)rel0 + lane;
            const uint64_t left = len - rel0;  // > 0
            const uint32_t span = left < kJumpTile ? (uint32_t)left : kJumpTile;
            const uint32_t n_own = lane < span ? (span - lane + 31) / 32 : 0;
            uint32_t pos[kJumpChains], dj[kJumpChains];
            uint32_t act = 0, nexti = 0;
#pragma unroll
            for (int c = 0; c < kJumpChains; c++) {
                pos[c] = base;
                dj[c] = 0;
                if (nexti < n_own) {
                    pos[c] = base + 32 * nexti++;
                    act |= 1u << c;
                }
            }
            while (act) {
                uint32_t e[kJumpChains];
#pragma unroll
                for (int c = 0; c < kJumpChains; c++) e[c] = jump_ld_if(cells + (pos[c] - dj[c]), act & (1u << c));
#pragma unroll
                for (int c = 0; c < kJumpChains; c++) {
                    if (!(act & (1u << c))) continue;
                    if (e[c]) {
                        if (dj[c]) {
                            dj[c] += e[c];
                            __stcg(cells + pos[c], dj[c]);
                        } else {
                            dj[c] = e[c];  // that was the byte's own cell
                        }
                    } else if (nexti < n_own) {  // a literal byte, or a walk that reached one: the cell holds the distance
                        pos[c] = base + 32 * nexti++;
                        dj[c] = 0;
                    } else {
                        act &= ~(1u << c);
                    }
                }
            }
            for (uint32_t i0 = 0; i0 < n_own; i0 += kJumpChains) {
                uint32_t d[kJumpChains], v[kJumpChains];
#pragma unroll
                for (int c = 0; c < kJumpChains; c++) d[c] = i0 + c < n_own ? __ldcg(cells + (base + 32 * (i0 + c))) : 0;
#pragma unroll
                for (int c = 0; c < kJumpChains; c++) v[c] = ld_u8_if(out + (base + 32 * (i0 + c) - d[c]), d[c]);
#pragma unroll
                for (int c = 0; c < kJumpChains; c++)
                    if (d[c]) out[base + 32 * (i0 + c)] = (uint8_t)v[c];
            }
        }
#else
        for (uint32_t sub = 0; sub < kJumpTile; sub += 32 * kJumpChains) {
            const uint32_t r0 = (uint32_t)rel0 + sub + lane;
            if (r0 - lane >= len) break;
            uint32_t dj[kJumpChains];
            uint32_t open = 0;
#pragma unroll
            for (int c = 0; c < kJumpChains; c++) {
                const uint32_t rel = r0 + c * 32;
                dj[c] = rel < len ? __ldcg(cells + rel) : 0;
                if (dj[c]) open |= 1u << c;
            }
            const uint32_t moved = open;  // the match bytes among mine
#if SZB_JUMP_BATCHED
            while (open) {
                uint32_t e[kJumpChains];
#pragma unroll
                for (int c = 0; c < kJumpChains; c++) e[c] = jump_ld_if(cells + (r0 + c * 32 - dj[c]), open & (1u << c));  // all in flight
#pragma unroll
                for (int c = 0; c < kJumpChains; c++) {
                    if (e[c]) {
                        dj[c] += e[c];
                        __stcg(cells + (r0 + c * 32), dj[c]);  // bytes that hang on this one skip what it has skipped
                    } else {
                        open &= ~(1u << c);  // reached a literal (or was not walking)
                    }
                }
            }
            uint32_t v[kJumpChains];
#pragma unroll
            for (int c = 0; c < kJumpChains; c++) v[c] = ld_u8_if(out + (r0 + c * 32 - dj[c]), moved & (1u << c));
#pragma unroll
            for (int c = 0; c < kJumpChains; c++)
                if (moved & (1u << c)) out[r0 + c * 32] = (uint8_t)v[c];
#else
            while (open) {
#pragma unroll
                for (int c = 0; c < kJumpChains; c++) {
                    if (open & (1u << c)) {
                        const uint32_t rel = r0 + c * 32;
                        const uint32_t e = jump_ld(cells + (rel - dj[c]));
                        if (e) {
                            dj[c] += e;
                            __stcg(cells + rel, dj[c]);  // bytes that hang on this one skip what it has skipped
                        } else {
                            open &= ~(1u << c);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kJumpChains; c++) {
                if (moved & (1u << c)) {
                    const uint32_t rel = r0 + c * 32;
                    out[rel] = out[rel - dj[c]];
                }
            }
#endif
        }
#endif
    }
}

// ---- k_long_verdict: one thread per long frame ----
__global__ void k_long_verdict(DeviceBatch a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.n_long) return;
    const unsigned long long e = a.long_err[slot];
    if (e == kLongNoError) return;
    const uint32_t f = a.exec_list[slot];
    a.frame_status[f] = -(int32_t)(e & 0xFF);
    a.frame_out_len[f] = 0;
}

