// execute.cuh -- stage 4 of the decode path (sm_100a): block positions, frame verdicts, block bodies, sequence execution,
// optional content checksums.  Included by kernels.cuh inside namespace szb (the kernel list is in kernels.cuh's header).
// Everything here is plain CUDA C++ apart from one prefetch instruction, so that tests/host_sim can run these kernels on an
// emulated CTA (tests/host_sim/warpsim.h) in the CPU suite.
#pragma once

#if defined(__CUDACC__)
#define SZB_PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#else
#define SZB_PREFETCH_L1(p) ((void)(p))
#endif

// ---------------------------------------------------------------------------------------------
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

// Bulk copy by one warp for large bodies (Raw blocks, long literal tails): destination words are
// written whole (4-byte aligned), each assembled from the two aligned source words that cover it.
// src and dst must not overlap.  Reads only aligned words that contain at least one source byte.
__device__ __forceinline__ void warp_memcpy(uint8_t *dst, const uint8_t *src, uint64_t n, uint32_t lane) {
    if (n < 64) {
        warp_copy(dst, src, n, lane);
        return;
    }
    const uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3)) & 3;
    if (lane < head) dst[lane] = src[lane];
    dst += head;
    src += head;
    n -= head;
    const uint64_t words = n >> 2;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src - mis);
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst);
    uint64_t i = lane;
    if (mis == 0) {
        for (; i + 96 < words; i += 128) {  // four independent loads in flight per lane
            const uint32_t v0 = sw[i], v1 = sw[i + 32], v2 = sw[i + 64], v3 = sw[i + 96];
            dw[i] = v0;
            dw[i + 32] = v1;
            dw[i + 64] = v2;
            dw[i + 96] = v3;
        }
        for (; i < words; i += 32) dw[i] = sw[i];
    } else {
        for (; i + 96 < words; i += 128) {
            const uint32_t a0 = sw[i], a1 = sw[i + 1], b0 = sw[i + 32], b1 = sw[i + 33];
            const uint32_t c0 = sw[i + 64], c1 = sw[i + 65], d0 = sw[i + 96], d1 = sw[i + 97];
            dw[i] = __funnelshift_r(a0, a1, mis * 8);
            dw[i + 32] = __funnelshift_r(b0, b1, mis * 8);
            dw[i + 64] = __funnelshift_r(c0, c1, mis * 8);
            dw[i + 96] = __funnelshift_r(d0, d1, mis * 8);
        }
        for (; i < words; i += 32) dw[i] = __funnelshift_r(sw[i], sw[i + 1], mis * 8);
    }
    const uint32_t tail = (uint32_t)(n & 3);
    if (lane < tail) dst[(words << 2) + lane] = src[(words << 2) + lane];
}
__device__ __forceinline__ void warp_memset(uint8_t *dst, uint8_t v, uint64_t n, uint32_t lane) {
    if (n < 64) {
        warp_fill(dst, v, n, lane);
        return;
    }
    const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (lane < head) dst[lane] = v;
    dst += head;
    n -= head;
    const uint32_t w = v * 0x01010101u;
    const uint4 q = make_uint4(w, w, w, w);
    uint4 *dq = reinterpret_cast<uint4 *>(dst);
    const uint64_t quads = n >> 4;
    for (uint64_t i = lane; i < quads; i += 32) dq[i] = q;
    const uint32_t tail = (uint32_t)(n & 15);
    if (lane < tail) dst[(quads << 4) + lane] = v;
}

// ---------------------------------------------------------------------------------------------
// Stage 4.  The reference pushes every literal run and every match through a window ring
// buffer (ringbuffer.go:102-277).  Here the whole output lives in HBM and the "window" is just
// earlier output.  Execution is split in two halves that meet in shared memory:
//
//   * the PRODUCER side takes 32 sequences per round, resolves what is sequential about them with
//     warp scans (offsets through the repeat history, positions as prefix sums) and appends
//     their SEGMENTS -- a literal run or a match, each a contiguous piece of output with a
//     contiguous source -- to a ring: one 64-bit word per segment (source address minus output
//     position, so source = word + position for every byte of it) plus one bit in a position bitmap;
//   * the CONSUMER side produces the output in address order, one aligned 128-byte line per step,
//     lane i making bytes i, 32+i, 64+i, 96+i of the line.  A byte finds its segment with a
//     popcount over the bitmap, loads its source byte, and the line leaves as one word per lane.
//     Neighbouring lanes read neighbouring bytes, so a step touches a handful of cache lines.
//     A source below the line is already in memory (written by an earlier step, block or kernel)
//     and is read back through L1/L2; a source inside the line (offset < 128) is another byte of
//     the step: earlier 32-byte chunks are exchanged through shared memory, the own chunk by
//     pointer jumping over shuffles.
constexpr uint32_t kRingBits = 4096;              // output positions the bitmap covers
constexpr uint32_t kSpanBytes = kRingBits - 256;  // a round may reach this far past the line being consumed
constexpr uint32_t kSegRing = 256;                // >= 2 x 64 segments of two rounds + the segments of a partial line (<= 64) + 1
constexpr uint32_t kConstRun = 256;               // RLE literal runs up to this long are segments (their source is a row of DeviceBatch::bytefill)
struct ExecSmem {
    unsigned long long seg[kSegRing];             // per segment: source address minus output position
    __align__(16) uint32_t bits[kRingBits / 32];  // bit p % kRingBits set: a segment starts at output position p
    __align__(16) uint8_t grp[128];               // the bytes of the step in flight
};

struct ExecState {  // the consumer's side
    uint64_t line;  // next line to produce (multiple of 128)
    uint32_t head;  // output below line + head is in memory already (0 .. 128)
    uint32_t seen;  // segments that start below line + head
};

// Produces the bytes [lo, hi) of the line at st.line (positions relative to the line; kWhole: all 128).
template <bool kWhole>
__device__ __forceinline__ void place_step(ExecSmem &sm, uint8_t *dst, ExecState &st, uint32_t lo, uint32_t hi, uint32_t lane,
                                           uint32_t le_mask) {
    uint32_t *bw = &sm.bits[(uint32_t)(st.line >> 5) & (kRingBits / 32 - 1)];
    const uint4 m4 = *reinterpret_cast<const uint4 *>(bw);
    const uint32_t m[4] = {m4.x, m4.y, m4.z, m4.w};
    const uint64_t my_pos = st.line + lane;  // output position of my byte of chunk 0
    const uint64_t low_addr = reinterpret_cast<uintptr_t>(dst) + st.line + lo;  // a source at or above this address is a byte of this step
    const uint32_t low_lo = (uint32_t)low_addr, low_hi = (uint32_t)(low_addr >> 32);
    uint32_t last = st.seen - 1;  // the last segment that starts below the chunk
    uint32_t v[4], sg[4];
    bool ing[4];
    // every byte of the line finds its segment and issues its load
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t rel = (c << 5) + lane;
        const uint32_t ord = last + __popc(m[c] & le_mask);  // the last segment that starts at or before my byte
        last += __popc(m[c]);
        const bool live = kWhole || (rel >= lo && rel < hi);
        const uint64_t src = sm.seg[ord & (kSegRing - 1)] + my_pos + (c << 5);
        // Sources in [low_addr, my own address) repeat a byte of this step that is not in memory yet.  Only a match can
        // point there, and a line does not straddle a 4 GiB boundary: compare the low words, then the high ones.
        sg[c] = (uint32_t)src - low_lo;
        ing[c] = live && sg[c] < rel - lo && (uint32_t)(src >> 32) == low_hi;
        sg[c] += lo;  // position in the line
        v[c] = 0;
        if (live && !ing[c]) v[c] = *reinterpret_cast<const uint8_t *>(src);
    }
    st.seen = last + 1;
    uint8_t *const my_grp = sm.grp + lane;
    if (!__any_sync(kFull, ing[0] | ing[1] | ing[2] | ing[3])) {
#pragma unroll
        for (int c = 0; c < 4; c++) my_grp[c << 5] = (uint8_t)v[c];
    } else {
        // bytes that repeat bytes of this step: earlier chunks through shared memory, the own chunk by
        // pointer jumping over shuffles (chains of in-chunk sources halve every round)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t open = __ballot_sync(kFull, ing[c]);
            if (open) {
                bool unres = ing[c];
                if (c > 0) {
                    __syncwarp();  // the chunks before this one are in sm.grp
                    if (unres && (sg[c] >> 5) < (uint32_t)c) {
                        v[c] = sm.grp[sg[c]];
                        unres = false;
                    }
                    open = __ballot_sync(kFull, unres);
                }
                uint32_t par = sg[c] & 31;
                while (open) {
                    const uint32_t pv = __shfl_sync(kFull, v[c], par);
                    const uint32_t pp = __shfl_sync(kFull, par, par);
                    if (unres) {
                        if (!((open >> par) & 1)) {
                            v[c] = pv;
                            unres = false;
                        } else {
                            par = pp;
                        }
                    }
                    open = __ballot_sync(kFull, unres);
                }
            }
            my_grp[c << 5] = (uint8_t)v[c];
        }
    }
    __syncwarp();
    // out: one aligned word per lane, a full line per warp; single bytes where the line is partial
    const uint32_t wrel = lane << 2;
    const uint32_t word = reinterpret_cast<const uint32_t *>(sm.grp)[lane];
    uint8_t *out = dst + st.line;
    if (kWhole || (wrel >= lo && wrel + 4 <= hi)) {
        *reinterpret_cast<uint32_t *>(out + wrel) = word;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (wrel + k >= lo && wrel + k < hi) out[wrel + k] = (uint8_t)(word >> (8 * k));
    }
    if (lane < 4) bw[lane] = 0;  // the bitmap is a ring: leave it clean for the next lap
    __syncwarp();
}

// produce every complete line below the output position `limit` (segments must cover the output up to there)
__device__ __forceinline__ void exec_drain(ExecSmem &sm, uint8_t *dst, ExecState &st, uint64_t limit, uint32_t lane, uint32_t le_mask) {
    uint32_t n = ((uint32_t)limit - (uint32_t)st.line) >> 7;  // limit - line < 2^32: the low words do
    if (n == 0) return;
    if (st.head) {  // the rest of a line that was flushed in part
        place_step<false>(sm, dst, st, st.head, 128, lane, le_mask);
        st.line += 128;
        st.head = 0;
        n--;
    }
    for (; n; n--) {
        place_step<true>(sm, dst, st, 0, 128, lane, le_mask);
        st.line += 128;
    }
}
// produce everything up to `prod`, the last, partial line included: everything below prod is then in memory
__device__ __forceinline__ void exec_flush(ExecSmem &sm, uint8_t *dst, ExecState &st, uint64_t prod, uint32_t lane, uint32_t le_mask) {
    exec_drain(sm, dst, st, prod, lane, le_mask);
    const uint32_t hi = (uint32_t)(prod - st.line);  // < 128
    if (hi > st.head) {
        place_step<false>(sm, dst, st, st.head, hi, lane, le_mask);
        st.head = hi;
    }
}
// continue at another output position (everything flushed)
__device__ __forceinline__ void exec_seek(ExecState &st, uint64_t pos) {
    st.line = pos & ~(uint64_t)127;
    st.head = (uint32_t)pos & 127;
}

// Where the producer's segments go.  InlineSink: the producing warp is the consumer too, and stays one append behind
// so that the prefetches it issued have time to land.
struct InlineSink {
    ExecSmem &sm;
    uint8_t *dst;
    ExecState st;
    uint32_t lane, le_mask;
    __device__ __forceinline__ uint64_t line() const { return st.line; }
    __device__ __forceinline__ void appended(uint64_t prev_prod, uint64_t) { exec_drain(sm, dst, st, prev_prod, lane, le_mask); }
    __device__ __forceinline__ void drain(uint64_t prod) { exec_drain(sm, dst, st, prod, lane, le_mask); }
    __device__ __forceinline__ void flush(uint64_t prod) { exec_flush(sm, dst, st, prod, lane, le_mask); }
    __device__ __forceinline__ void seek(uint64_t pos) { exec_seek(st, pos); }
    __device__ __forceinline__ void finish(uint64_t prod) { exec_flush(sm, dst, st, prod, lane, le_mask); }
};

// PairSink: a second warp of the CTA consumes (k_execute_pair).  The two warps meet at one __syncthreads per command;
// a command is executed by the consumer while the producer works on the next round.
enum : uint32_t { kCmdNop = 0, kCmdDrain = 1, kCmdFlush = 2, kCmdSeek = 3, kCmdExit = 4 };
struct PairShared {
    unsigned long long line[2];  // the consumer's st.line after command i, in slot i & 1
    unsigned long long arg[2];
    uint32_t cmd[2];
};
struct PairSink {
    PairShared &sh;
    uint32_t lane, it;
    uint64_t seen_line;  // what the consumer had reached one command ago: a lower bound, which is all the producer needs
    __device__ __forceinline__ uint64_t line() const { return seen_line; }
    __device__ __forceinline__ void hand(uint32_t cmd, uint64_t arg) {
        if (lane == 0) {
            sh.cmd[it & 1] = cmd;
            sh.arg[it & 1] = arg;
        }
        __syncthreads();  // command `it` starts; command it-1 is complete and published its line before this barrier
        if (it) seen_line = sh.line[(it - 1) & 1];
        it++;
    }
    __device__ __forceinline__ void appended(uint64_t, uint64_t prod) { hand(kCmdDrain, prod); }
    __device__ __forceinline__ void drain(uint64_t prod) { hand(kCmdDrain, prod); }
    __device__ __forceinline__ void flush(uint64_t prod) {  // returns when everything below prod is in memory
        hand(kCmdFlush, prod);
        hand(kCmdNop, 0);
    }
    __device__ __forceinline__ void seek(uint64_t pos) { hand(kCmdSeek, pos); }
    __device__ __forceinline__ void finish(uint64_t prod) {
        hand(kCmdFlush, prod);
        hand(kCmdExit, 0);
    }
};

// One warp per frame, before any output is written: the frame's verdict (the first failing
// block decides, as in the sequential reference; then the header walk's verdict; then capacity)
// and its placement in dst.
__global__ void k_frame_verdict(DeviceBatch a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per frame: a frame may have 10^5 blocks
    if (f >= a.nframes) return;
    const szb_frame_desc fr = a.frames[f];
    const uint32_t b0 = fr.first_block, nb = fr.nblocks;
    int err = SZB_OK;
    uint32_t nexec = nb;  // leading blocks that are fine as far as stages 1-3 and the header walk can tell
    // k_resolve walked the frame in order (place.cuh): its verdict covers the blocks' entropy stages and the execution
    const int ps = a.place_state ? a.place_state[f] : 1;
    if (ps <= 0) {
        err = ps;
        nexec = ps == 0 ? nb : 0;  // 0: k_place executes the whole frame; < 0: the error k_resolve found
    } else {
        for (uint32_t i0 = 0; i0 < nb; i0 += 32) {
            int e = SZB_OK;
            if (i0 + lane < nb) {
                const int ls = a.lit_status[b0 + i0 + lane], ss = a.seq_status[b0 + i0 + lane];
                const int hs = a.blocks[b0 + i0 + lane].hdr_status;  // the walk's verdict on this block's sequences header
                if ((ls | ss | hs) != 0) e = ls ? ls : (ss ? ss : hs);
            }
            const uint32_t bad = __ballot_sync(kFull, e != 0);
            if (bad) {
                const uint32_t first = (uint32_t)__ffs(bad) - 1;
                err = __shfl_sync(kFull, e, first);
                nexec = i0 + first;
                break;
            }
        }
    }
    if (lane != 0) return;
    if (err == SZB_OK) err = fr.status;  // blocks after a failing header are absent from the table
    if (a.total[0] > a.dst_cap) {
        if (err == SZB_OK) err = SZB_ERR_DST_TOO_SMALL;
        nexec = 0;
    }
    const uint64_t frame_base = nb ? a.out_off[b0] : 0;
    if (a.frame_nexec) a.frame_nexec[f] = nexec;
    a.frame_status[f] = err;
    a.frame_out_off[f] = frame_base;
    a.frame_out_len[f] = (nb && err == SZB_OK) ? a.out_off[b0 + nb - 1] + a.out_size[b0 + nb - 1] - frame_base : 0;
}

// Blocks whose output does not depend on earlier output -- Raw bodies (framedecompressor.go:211-215),
// RLE bodies (:229-241) and compressed blocks without sequences, whose output is their literals
// (sequences.go:395-400, sequence_execution.go:55-60) -- are written by one warp each, all in
// parallel, before the per-frame sequence execution starts.
__global__ void __launch_bounds__(kCtaThreads) k_execute_bodies(DeviceBatch a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (w >= a.n_body) return;
    const uint32_t b = a.body_list[w];
    const szb_block_desc d = a.blocks[b];
    if (a.frame_nexec ? b - a.frames[d.frame].first_block >= a.frame_nexec[d.frame] : a.frame_status[d.frame] != SZB_OK) return;
    uint8_t *out = a.dst + a.out_off[b];
    const uint8_t *payload = a.src + d.src_off;
    if (d.type == 0) {
        warp_memcpy(out, payload, d.block_size, lane);
    } else if (d.type == 1) {
        warp_memset(out, payload[0], d.block_size, lane);
    } else if (d.lit_type == 1) {
        warp_memset(out, payload[d.lit_hdr_bytes], d.lit_regen, lane);
    } else {
        warp_memcpy(out, d.lit_type == 0 ? payload + d.lit_hdr_bytes : a.litbuf + d.lit_buf_off, d.lit_regen, lane);
    }
}

// The producer half for one frame (status OK), one warp: blocks in order, 32 sequences per round
// (sequence_execution.go:14-63); segments go to the ring in `sm`, the sink decides who consumes them and when.
template <class Sink>
__device__ __forceinline__ void produce_frame(const DeviceBatch &a, uint32_t f, ExecSmem &sm, Sink &sink, uint32_t lane) {
    const szb_frame_desc fr = a.frames[f];
    const uint32_t b0 = fr.first_block, nb = a.frame_nexec ? a.frame_nexec[f] : fr.nblocks;  // k_frame_verdict
    uint8_t *const dst = a.dst;
    int err = SZB_OK;
    const uint64_t frame_base = nb ? a.out_off[b0] : 0;
    const uint32_t lt_mask = 0x7FFFFFFFu >> (31 - lane);  // bits 0..lane-1
    uint64_t prod = frame_base;  // segments cover the output up to here
    uint32_t nseg = 0;           // segments appended so far

    History hist{1, 4, 8};  // framedecompressor.go:48,59
    for (uint32_t bi = 0; bi < nb && err == SZB_OK; bi++) {
        const uint32_t b = b0 + bi;
        const szb_block_desc d = a.blocks[b];
        if (d.type != 2 || d.nseq == 0) continue;  // written by k_execute_bodies already
        const uint8_t *payload = a.src + d.src_off;
        uint64_t out_pos = a.out_off[b];
        if (out_pos != prod) {  // blocks in between were written elsewhere
            sink.flush(prod);
            sink.seek(out_pos);
            prod = out_pos;
        }
        // Compressed: ExecuteSequences (sequence_execution.go:14-63)
        const bool lit_rle = d.lit_type == 1;
        // RLE literals: every literal byte is payload[lit_hdr_bytes]; runs read it from that byte's row of the fill table
        const uint8_t *__restrict__ lit = lit_rle ? a.bytefill + 256 * (uint32_t)payload[d.lit_hdr_bytes]
                                                  : (d.lit_type == 0 ? payload + d.lit_hdr_bytes : a.litbuf + d.lit_buf_off);
        const uint32_t nseq = d.nseq;
        const uint64_t sbo = d.seq_buf_off;
        uint32_t lit_pos = 0;
        // the triples are prefetched to L1 two rounds ahead (one line per array and round): lanes 0..2 take one array each
        const uint32_t *const my_seq = a.seq_ll + sbo + (lane < 3 ? lane : 0) * a.seq_stride;
        if (lane < 3) {
            SZB_PREFETCH_L1(my_seq);
            if (nseq > 32) SZB_PREFETCH_L1(my_seq + 32);
        }
        for (uint32_t base = 0; base < nseq; base += 32) {
            const uint32_t cnt = nseq - base < 32 ? nseq - base : 32;
            const bool act = lane < cnt;
            const uint32_t *const tr = a.seq_ll + (sbo + base + lane);  // the arrays are padded to whole rounds: no bounds needed
            uint32_t ll = tr[0], ml = tr[a.seq_stride], ofv = tr[2 * a.seq_stride];
            if (!act) {
                ll = 0;
                ml = 0;
                ofv = 4;
            }
            if (!lit_rle && lane == 0) SZB_PREFETCH_L1(lit + lit_pos + 256);
            if (lane < 3 && base + 64 < nseq) SZB_PREFETCH_L1(my_seq + base + 64);

            // --- offsets through the 3-entry history (nextOffset) ---
            // A sequence with a direct offset (offset value > 3) pushes it onto the history whatever the history
            // holds; only the repeat codes look at it.  So the walk jumps from repeat code to repeat code: the run of
            // direct sequences in between is folded in at once (its last three offsets are the new history).
            uint32_t off = ofv - 3;
            {
                uint32_t rm = __ballot_sync(kFull, act && ofv <= 3);
                uint32_t p = 0;  // sequences [0, p) are folded into hist
                for (;;) {
                    const uint32_t j = rm ? (uint32_t)__ffs(rm) - 1 : cnt;  // the next repeat code, or the end of the round
                    const uint32_t n = j - p;
                    if (n) {
                        const uint32_t o1 = __shfl_sync(kFull, off, j - 1);
                        const uint32_t o2 = __shfl_sync(kFull, off, n >= 2 ? j - 2 : 0);
                        const uint32_t o3 = __shfl_sync(kFull, off, n >= 3 ? j - 3 : 0);
                        if (n >= 3) {
                            hist = History{o1, o2, o3};
                        } else if (n == 2) {
                            hist = History{o1, o2, hist.h0};
                        } else {
                            hist = History{o1, hist.h0, hist.h1};
                        }
                    }
                    if (j >= cnt) break;
                    const uint32_t v = __shfl_sync(kFull, ofv, j);
                    const uint32_t l = __shfl_sync(kFull, ll, j);
                    const uint32_t o = next_offset(hist, v, l == 0);  // every lane tracks the same history
                    if (lane == j) off = o;
                    rm &= rm - 1;
                    p = j + 1;
                }
            }

            // --- positions: prefix sums over the round ---
            const uint32_t tot = ll + ml;
            uint32_t incl_ll, incl_tot;
            if (__reduce_max_sync(kFull, tot) < 2048) {
                // both sums stay below 2^16: one scan over (literals | literals + match << 16)
                uint32_t x = ll | (tot << 16);
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, x, dlt);
                    if ((int)lane >= dlt) x += t;
                }
                incl_ll = x & 0xFFFF;
                incl_tot = x >> 16;
            } else {
                incl_ll = ll;
                incl_tot = tot;
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                    const uint32_t t1 = __shfl_up_sync(kFull, incl_ll, dlt);
                    const uint32_t t2 = __shfl_up_sync(kFull, incl_tot, dlt);
                    if ((int)lane >= dlt) {
                        incl_ll += t1;
                        incl_tot += t2;
                    }
                }
            }
            const uint32_t round_ll = __shfl_sync(kFull, incl_ll, 31);
            const uint32_t round_tot = __shfl_sync(kFull, incl_tot, 31);
            if ((uint64_t)lit_pos + round_ll > d.lit_regen) {
                // literals.go:398-409 Read runs dry / sequence_execution.go:26-28; RLE literals: GetRest panics
                err = lit_rle ? SZB_ERR_PANIC : SZB_ERR_DIDNT_COPY_ALL_LITERAL_BYTES;
                break;
            }
            const uint32_t excl_tot = incl_tot - tot;  // my literal run starts at out_pos + excl_tot
            const uint32_t excl_ll = incl_ll - ll;     // and reads the literals from lit_pos + excl_ll
            {
                // every match must lie inside the frame (ringbuffer.go:203-214); a match length of 0 cannot come out of
                // stage 3 (ML codes start at 3, predefined.go:36-50) and would break the segment count below
                const uint64_t fb = out_pos - frame_base;  // frame bytes in front of the round
                bool bad;
                if (fb < 0x7F000000u) {
                    bad = off > (uint32_t)fb + excl_tot + ll;
                } else {
                    bad = off > fb + excl_tot + ll;
                }
                if (__any_sync(kFull, act && (bad || off == 0 || ml == 0))) {
                    err = SZB_ERR_CANT_REPEAT_BYTES;
                    break;
                }
            }

            // --- the round's segments go to the ring, as many sequences at a time as the bitmap holds (normally all) ---
            uint32_t start = 0;
            while (start < cnt) {
                const uint64_t line = sink.line();
                const uint32_t out_rel = (uint32_t)out_pos - (uint32_t)line;  // < 2^32: the low words do
                const uint32_t my_rel = out_rel + excl_tot;  // my literal run, relative to the line being consumed
                uint32_t nfit;
                if (start == 0 && !lit_rle && out_rel + round_tot <= kSpanBytes) {
                    nfit = cnt;  // the usual case: the whole round fits the ring
                } else {
                    const bool fits = my_rel + tot <= kSpanBytes && !(lit_rle && ll > kConstRun);
                    const uint32_t fitmask = __ballot_sync(kFull, fits && lane < cnt) >> start;
                    nfit = fitmask == (0xFFFFFFFFu >> start) ? 32 - start : __ffs(~fitmask) - 1;  // leading fits
                }
                if (nfit == 0) {
                    // one sequence longer than the ring: the whole warp on its literals, then on its match
                    sink.flush(prod);
                    const uint32_t L = __shfl_sync(kFull, ll, start), ML = __shfl_sync(kFull, ml, start);
                    const uint32_t OFF = __shfl_sync(kFull, off, start);
                    const uint64_t D = out_pos + __shfl_sync(kFull, excl_tot, start);
                    if (lit_rle)
                        warp_memset(dst + D, lit[0], L, lane);
                    else
                        warp_memcpy(dst + D, lit + lit_pos + __shfl_sync(kFull, excl_ll, start), L, lane);
                    __syncwarp();
                    uint8_t *MD = dst + D + L;
                    const uint8_t *MS = MD - OFF;
                    if (OFF >= 32) {
                        for (uint32_t k0 = 0; k0 < ML; k0 += 32) {
                            const uint32_t k = k0 + lane;
                            if (k < ML) MD[k] = MS[k];
                            __syncwarp();
                        }
                    } else if (ML) {  // overlapping: periodic extension of the OFF bytes before the match
                        for (uint32_t k = lane; k < ML; k += 32) MD[k] = MS[k % OFF];
                    }
                    __syncwarp();
                    prod = D + L + ML;
                    sink.seek(prod);
                    start++;
                    continue;
                }
                const uint32_t end = start + nfit;
                const bool in = lane >= start && lane < end;
                const uint32_t no_lit = __ballot_sync(kFull, in && ll == 0);
                if (in) {
                    uint32_t ord = nseg + 2 * (lane - start) - __popc(no_lit & lt_mask);
                    const uint32_t bit0 = ((uint32_t)line & (kRingBits - 1)) + my_rel;  // my literal run in the bitmap
                    if (ll) {
                        // literal byte at output position p: lit[lit_pos + excl_ll + (p - my start)]; a run of RLE literals reads
                        // the first bytes of the fill row
                        const uint64_t my_start = out_pos + excl_tot;
                        sm.seg[ord & (kSegRing - 1)] = reinterpret_cast<uintptr_t>(lit) + (lit_rle ? 0 : lit_pos + excl_ll) - my_start;
                        atomicOr(&sm.bits[(bit0 >> 5) & (kRingBits / 32 - 1)], 1u << (bit0 & 31));
                        ord++;
                    }
                    {
                        const uint32_t bit1 = bit0 + ll;
                        sm.seg[ord & (kSegRing - 1)] = reinterpret_cast<uintptr_t>(dst) - off;
                        atomicOr(&sm.bits[(bit1 >> 5) & (kRingBits / 32 - 1)], 1u << (bit1 & 31));
                        // the consumer gets here about a round later: have the source on its way to L1
                        SZB_PREFETCH_L1(dst + (out_pos + excl_tot + ll - off));
                    }
                }
                nseg += 2 * nfit - __popc(no_lit);
                const uint64_t prev_prod = prod;
                prod = line + __shfl_sync(kFull, my_rel + tot, end - 1);
                __syncwarp();
                sink.appended(prev_prod, prod);
                start = end;
            }
            out_pos += round_tot;
            lit_pos += round_ll;
        }
        if (err != SZB_OK) break;
        // trailing literals (sequence_execution.go:55-60, literals.go:411-420): one more segment, or a bulk copy
        const uint32_t rest = d.lit_regen - lit_pos;
        if (rest && (prod - sink.line()) + rest <= kSpanBytes && !(lit_rle && rest > kConstRun)) {
            if (lane == 0) {
                const uint32_t bit0 = (uint32_t)prod & (kRingBits - 1);
                sm.seg[nseg & (kSegRing - 1)] = reinterpret_cast<uintptr_t>(lit) + (lit_rle ? 0 : lit_pos) - out_pos;
                atomicOr(&sm.bits[bit0 >> 5], 1u << (bit0 & 31));
            }
            nseg++;
            prod += rest;
            __syncwarp();
            sink.drain(prod);
        } else if (rest) {
            sink.flush(prod);
            if (lit_rle)
                warp_memset(dst + out_pos, lit[0], rest, lane);
            else
                warp_memcpy(dst + out_pos, lit + lit_pos, rest, lane);
            __syncwarp();
            prod = out_pos + rest;
            sink.seek(prod);
        }
    }
    // On an error the frame's output is void; what the ring still holds is written anyway (it is within the frame's range).
    sink.finish(prod);
    if (lane == 0 && err != SZB_OK) {  // failed while executing (bad offset, literals ran dry)
        a.frame_status[f] = err;
        a.frame_out_len[f] = 0;
    }
}

#include "execute_long.cuh"
#include "place.cuh"
#include "exec2.cuh"

// One warp per frame: it produces the segments and consumes them.  Frames exec_list[first_slot, first_slot + n_slots).
#ifndef SZB_EXEC_MIN_CTAS
#define SZB_EXEC_MIN_CTAS 8
#endif
__global__ void __launch_bounds__(kCtaThreads, SZB_EXEC_MIN_CTAS) k_execute(DeviceBatch a, uint32_t first_slot, uint32_t n_slots) {
    __shared__ ExecSmem smem[kWarpsPerCta];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t slot = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (slot >= n_slots) return;
    const uint32_t f = a.exec_list[first_slot + slot];
    if (a.frame_nexec ? a.frame_nexec[f] == 0 : a.frame_status[f] != SZB_OK) return;  // k_frame_verdict
    if (a.place_state && a.place_state[f] != 1) return;  // k_place executes it (place.cuh)
    if (x2_takes(a, f)) return;                          // k_execute2 executes it (exec2.cuh)
    if (a.frame_dict && a.frame_dict[f]) {               // only k_execute2 knows dictionaries (frames of 2 GiB and more: not with one)
        if (lane == 0) {
            a.frame_status[f] = SZB_ERR_UNSUPPORTED;
            a.frame_out_len[f] = 0;
        }
        return;
    }
    ExecSmem &sm = smem[threadIdx.x >> 5];
    for (uint32_t wd = lane; wd < kRingBits / 32; wd += 32) sm.bits[wd] = 0;
    __syncwarp();
    InlineSink sink{sm, a.dst, ExecState{}, lane, 0xFFFFFFFFu >> (31 - lane)};
    const szb_frame_desc fr = a.frames[f];
    exec_seek(sink.st, fr.nblocks ? a.out_off[fr.first_block] : 0);
    sink.st.seen = 0;
    produce_frame(a, f, sm, sink, lane);
}

// The few frames that are far longer than the rest finish last and then run almost alone: two warps per frame, one
// producing segments, one consuming them (PairSink).  Twice the warps per frame is the wrong trade while the SMs are full
// of frames, which is why only the longest frames take this path.
__global__ void __launch_bounds__(64) k_execute_pair(DeviceBatch a, uint32_t first_slot, uint32_t n_slots) {
    __shared__ ExecSmem sm;
    __shared__ PairShared sh;
    const uint32_t lane = threadIdx.x & 31;
    if (blockIdx.x >= n_slots) return;
    const uint32_t f = a.exec_list[first_slot + blockIdx.x];
    if (a.frame_status[f] != SZB_OK) return;  // k_frame_verdict; both warps agree
    if (long_jump_ok(a, first_slot + blockIdx.x)) return;  // taken by the block-parallel path (execute_long.cuh)
    if (a.pair2 && x2_takes(a, f)) return;                 // k_execute_pair2's (exec2.cuh)
    const szb_frame_desc fr = a.frames[f];
    const uint64_t frame_base = fr.nblocks ? a.out_off[fr.first_block] : 0;
    for (uint32_t wd = threadIdx.x; wd < kRingBits / 32; wd += 64) sm.bits[wd] = 0;
    __syncthreads();
    if (threadIdx.x < 32) {
        PairSink sink{sh, lane, 0, frame_base & ~(uint64_t)127};
        produce_frame(a, f, sm, sink, lane);
    } else {
        const uint32_t le_mask = 0xFFFFFFFFu >> (31 - lane);
        ExecState st;
        exec_seek(st, frame_base);
        st.seen = 0;
        for (uint32_t it = 0;; it++) {
            __syncthreads();
            const uint32_t cmd = sh.cmd[it & 1];
            const uint64_t arg = sh.arg[it & 1];
            if (cmd == kCmdExit) break;
            if (cmd == kCmdDrain)
                exec_drain(sm, a.dst, st, arg, lane, le_mask);
            else if (cmd == kCmdFlush)
                exec_flush(sm, a.dst, st, arg, lane, le_mask);
            else if (cmd == kCmdSeek)
                exec_seek(st, arg);
            if (lane == 0) sh.line[it & 1] = st.line;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Content checksum (NOT a reference behaviour: the reference leaves the 4 bytes
// unread, frame.go:105-108).  zstd stores the low 32 bits of XXH64(content, seed 0) after the last
// block.  XXH64 is a chain of four independent accumulator lanes over 32-byte stripes, so a frame
// offers no more than that; the batch offers one thread per frame.
__device__ __forceinline__ uint64_t xxh_rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t xxh_round(uint64_t acc, uint64_t in) {
    return xxh_rotl(acc + in * 0xC2B2AE3D27D4EB4FULL, 31) * 0x9E3779B185EBCA87ULL;
}
__device__ __forceinline__ uint64_t xxh_merge(uint64_t h, uint64_t v) {
    return (h ^ xxh_round(0, v)) * 0x9E3779B185EBCA87ULL + 0x85EBCA77C2B2AE63ULL;
}
// unaligned little-endian reads assembled from aligned words (only words holding a content byte are touched
// when n is what the caller needs)
__device__ __forceinline__ uint64_t xxh_read64(const uint8_t *p) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - mis);
    const uint32_t w0 = w[0], w1 = w[1], w2 = mis ? w[2] : 0u;
    return (uint64_t)__funnelshift_r(w0, w1, mis * 8) | ((uint64_t)__funnelshift_r(w1, w2, mis * 8) << 32);
}
__device__ __forceinline__ uint32_t xxh_read32(const uint8_t *p) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - mis);
    const uint32_t w0 = w[0], w1 = mis ? w[1] : 0u;
    return __funnelshift_r(w0, w1, mis * 8);
}
__device__ uint64_t xxh64(const uint8_t *p, uint64_t len) {
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL,
                   P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
        const uint8_t *limit = end - 32;
        do {
            v1 = xxh_round(v1, xxh_read64(p));
            v2 = xxh_round(v2, xxh_read64(p + 8));
            v3 = xxh_round(v3, xxh_read64(p + 16));
            v4 = xxh_round(v4, xxh_read64(p + 24));
            p += 32;
        } while (p <= limit);
        h = xxh_rotl(v1, 1) + xxh_rotl(v2, 7) + xxh_rotl(v3, 12) + xxh_rotl(v4, 18);
        h = xxh_merge(h, v1);
        h = xxh_merge(h, v2);
        h = xxh_merge(h, v3);
        h = xxh_merge(h, v4);
    } else {
        h = P5;
    }
    h += len;
    while (p + 8 <= end) {
        h ^= xxh_round(0, xxh_read64(p));
        h = xxh_rotl(h, 27) * P1 + P4;
        p += 8;
    }
    if (p + 4 <= end) {
        h ^= (uint64_t)xxh_read32(p) * P1;
        h = xxh_rotl(h, 23) * P2 + P3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t)(*p) * P5;
        h = xxh_rotl(h, 11) * P1;
        p++;
    }
    h ^= h >> 33;
    h *= P2;
    h ^= h >> 29;
    h *= P3;
    h ^= h >> 32;
    return h;
}

__global__ void k_verify_checksums(DeviceBatch a) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.nframes) return;
    const szb_frame_desc *fr = a.frames + f;
    if (!fr->checksum_valid || a.frame_status[f] != SZB_OK) return;  // no checksum, or its 4 bytes lay outside the caller's extent
    const uint64_t h = xxh64(a.dst + a.frame_out_off[f], a.frame_out_len[f]);
    if ((uint32_t)h != fr->checksum) {
        a.frame_status[f] = SZB_ERR_CHECKSUM_MISMATCH;
        a.frame_out_len[f] = 0;
    }
}

