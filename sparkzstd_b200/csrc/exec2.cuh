// exec2.cuh -- stage 4 for the frames one warp executes, second generation (sm_100a): k_execute2.
// Included by execute.cuh inside namespace szb.  Plain CUDA C++ (tests/host_sim runs it on an emulated CTA).
//
// replaces: decompression/sequence_execution.go:14-114 (ExecuteSequences, nextOffset) and ringbuffer.go:197-277 (match copy)
//
// The shape is k_execute's (execute.cuh): a producer half that turns 32 sequences per round into SEGMENTS (a literal run or a
// match: a contiguous piece of output with a contiguous source), and a consumer half that makes the output in address order,
// one aligned 128-byte line per step, lane i making bytes i, 32+i, 64+i, 96+i.  What changed is the arithmetic, because
// k_execute is bound by instruction issue (profiles/r02_*: 267 warp instructions per line, 170 of them in the consumer):
//
//   * positions are 32-bit "q positions", counted from the 128-byte aligned address at or below the frame's first byte,
//     so a line of output is a line of memory; a frame that regenerates 2 GiB or more stays with k_execute;
//   * a segment is ONE 32-bit entry: a match's offset (through the repeat history), or bit 31 | the difference between a
//     literal byte's q position and its index into the block's literals.  For both, `q - (entry & 0x7FFFFFFF)` is the index of
//     the source byte -- into the output for a match, into the literals for a literal run -- and "the source is a byte of
//     this very step" is one unsigned compare of the entry against the byte's place in the line (a literal entry is huge);
//   * the line leaves as four 32-byte stores (one byte per lane each), no transposition through shared memory; bytes that
//     repeat bytes of the same step take them from the lane that holds them: earlier 32-byte chunks with one shuffle of the
//     packed bytes, the own chunk by pointer jumping over shuffles;
//   * the block's literals have ONE base address, so the consumer is brought up to date at every block boundary (a partial
//     line per block).
//
// Shared memory per warp: 256 entries + a 4096-position bitmap = 1.5 KB (k_execute: 2.7 KB).
#pragma once

constexpr uint32_t kX2Bits = 4096;             // output positions the bitmap covers
constexpr uint32_t kX2Span = kX2Bits - 256;    // a round may reach this far past the line being consumed
constexpr uint32_t kX2Ring = 256;              // >= 2 x 64 segments of two rounds + the segments of a partial line (<= 64) + 1
constexpr uint32_t kX2ConstRun = 256;          // RLE literal runs up to this long are segments (source: a row of DeviceBatch::bytefill)
constexpr uint32_t kX2Lit = 0x80000000u;
constexpr uint64_t kX2MaxFrame = 0x7FFF0000ull;  // frames that regenerate this much or more are k_execute's

struct X2Smem {
    uint32_t seg[kX2Ring];                         // per segment: match offset, or kX2Lit | (q position - literal index)
    __align__(16) uint32_t bits[kX2Bits / 32];     // bit q % kX2Bits set: a segment starts at q
};

// Staged rounds (x2_round_staged below): the output of one round of 32 short sequences is assembled in shared memory by one
// LANE per sequence and leaves as whole 16-byte units.
#ifndef SZB_X2_STAGED
#define SZB_X2_STAGED 0  // measured (profiles/r02q_*): 13.1 ms against 8.97 ms on configs[1] -- 9.75 G warp instructions, not fewer
#endif
constexpr uint32_t kS2Cap = 1280;     // output bytes of a round that is staged (32 sequences of text: ~350)
constexpr uint32_t kS2Lit = 16;       // a lane copies its own literal run up to this length ...
constexpr uint32_t kS2Ml = 17;        // ... and its own match up to this length: what two aligned 16-byte loads cover at any alignment
#ifndef SZB_S2_MAX_DEFERRED
#define SZB_S2_MAX_DEFERRED 12
#endif
constexpr uint32_t kS2MaxDeferred = SZB_S2_MAX_DEFERRED;  // with more sequences than this left to the whole warp, the segment path is cheaper
constexpr uint32_t kS2ScrStride = 80;  // 2 x 16 literal bytes + 2 x 16 match-source bytes (+ 16: rows start in different banks)
struct X2Stage {
    __align__(16) uint8_t stg[kS2Cap + 48];        // [16 * u, 16 * u + 16) = the round's output unit u, counted from qpos & ~15
    __align__(16) uint8_t scr[32 * kS2ScrStride];  // per lane: the aligned 16-byte chunks that hold its literals and its match's source
};

struct X2State {
    uint8_t *qb;         // q position 0
    const uint8_t *lit;  // the literals of the block being executed
    uint32_t line;       // next line to produce (q position, multiple of 128)
    uint32_t head;       // output below line + head is in memory already (0 .. 128)
    uint32_t seen;       // segments that start below line + head
};

// every kernel evaluates the same predicate: does k_execute2 take frame f (status OK)?
__device__ __forceinline__ bool x2_takes(const DeviceBatch &a, uint32_t f) { return a.exec2 != 0 && a.frame_out_len[f] < kX2MaxFrame; }

// Produces the bytes [lo, hi) of the line at st.line (positions relative to the line; kWhole: all 128).
template <bool kWhole>
__device__ __forceinline__ void x2_step(X2Smem &sm, X2State &st, uint32_t lo, uint32_t hi, uint32_t lane, uint32_t le_mask) {
    uint32_t *bw = &sm.bits[(st.line >> 5) & (kX2Bits / 32 - 1)];
    const uint4 m4 = *reinterpret_cast<const uint4 *>(bw);
    const uint32_t m[4] = {m4.x, m4.y, m4.z, m4.w};
    const uint32_t q0 = st.line + lane;  // q position of my byte of chunk 0
    uint32_t last = st.seen - 1;         // the last segment that starts below the chunk
    uint32_t v[4], so[4];
    bool ing[4];
    // every byte of the line finds its segment and issues its load
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t rel = (c << 5) + lane;
        const uint32_t ord = last + __popc(m[c] & le_mask);  // the last segment that starts at or before my byte
        last += __popc(m[c]);
        const uint32_t e = sm.seg[ord & (kX2Ring - 1)];
        const bool live = kWhole || (rel >= lo && rel < hi);
        const uint32_t idx = q0 + (c << 5) - (e & ~kX2Lit);
        // A match whose offset does not reach below the step's first byte repeats a byte of this step that is not in memory yet.
        ing[c] = live && e <= rel - lo;
        so[c] = rel - e;  // its place in the line
        const uint8_t *p = ((int32_t)e < 0 ? st.lit : st.qb) + idx;
        v[c] = 0;
        if (live && !ing[c]) v[c] = *p;
    }
    st.seen = last + 1;
    if (__any_sync(kFull, ing[0] | ing[1] | ing[2] | ing[3])) {
        uint32_t W = 0;  // the bytes of the chunks done so far, chunk k in bits 8k..8k+7
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t open = __ballot_sync(kFull, ing[c]);
            if (open) {
                bool unres = ing[c];
                uint32_t par = so[c] & 31;
                if (c > 0) {
                    const uint32_t t = __shfl_sync(kFull, W, par);
                    if (unres && (so[c] >> 5) < (uint32_t)c) {
                        v[c] = (t >> ((so[c] >> 5) << 3)) & 0xFF;
                        unres = false;
                    }
                    open = __ballot_sync(kFull, unres);
                }
                while (open) {  // the own chunk: chains of in-chunk sources halve every round
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
            W |= v[c] << (c << 3);
        }
    }
    // out: four 32-byte rows, one byte per lane each
    uint8_t *out = st.qb + q0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const uint32_t rel = (c << 5) + lane;
        if (kWhole || (rel >= lo && rel < hi)) out[c << 5] = (uint8_t)v[c];
    }
    __syncwarp();                // every lane has read the line's bitmap words (the votes above converge the warp, but they
                                 // are no memory fence: racecheck, profiles/r02j_racecheck.txt)
    if (lane < 4) bw[lane] = 0;  // the bitmap is a ring: leave it clean for the next lap
    __syncwarp();
}

// produce every complete line below the q position `limit` (segments must cover the output up to there)
__device__ __forceinline__ void x2_drain(X2Smem &sm, X2State &st, uint32_t limit, uint32_t lane, uint32_t le_mask) {
    uint32_t n = (limit - st.line) >> 7;
    if (n == 0) return;
    if (st.head) {  // the rest of a line that was flushed in part
        x2_step<false>(sm, st, st.head, 128, lane, le_mask);
        st.line += 128;
        st.head = 0;
        n--;
    }
    for (; n; n--) {
        x2_step<true>(sm, st, 0, 128, lane, le_mask);
        st.line += 128;
    }
}
// produce everything up to `prod`, the last, partial line included: everything below prod is then in memory
__device__ __forceinline__ void x2_flush(X2Smem &sm, X2State &st, uint32_t prod, uint32_t lane, uint32_t le_mask) {
    x2_drain(sm, st, prod, lane, le_mask);
    const uint32_t hi = prod - st.line;  // < 128
    if (hi > st.head) {
        x2_step<false>(sm, st, st.head, hi, lane, le_mask);
        st.head = hi;
    }
}
// The same out of line, for the places a frame passes rarely (block starts, sequences longer than the ring, the frame's end):
// one inlined copy of the step per hot call site keeps the kernel inside the instruction cache.
__device__ __noinline__ X2State x2_drain_cold(X2Smem &sm, X2State st, uint32_t limit, uint32_t lane, uint32_t le_mask) {
    x2_drain(sm, st, limit, lane, le_mask);
    return st;
}
__device__ __noinline__ X2State x2_flush_cold(X2Smem &sm, X2State st, uint32_t prod, uint32_t lane, uint32_t le_mask) {
    st = x2_drain_cold(sm, st, prod, lane, le_mask);
    const uint32_t hi = prod - st.line;  // < 128
    if (hi > st.head) {
        x2_step<false>(sm, st, st.head, hi, lane, le_mask);
        st.head = hi;
    }
    return st;
}
// continue at another q position (everything flushed)
__device__ __forceinline__ void x2_seek(X2State &st, uint32_t q) {
    st.line = q & ~127u;
    st.head = q & 127u;
}

// One round of 32 sequences whose output fits the staging buffer, everything below qpos in memory (flushed).
//
// The per-byte gather of x2_step costs ~16 warp instructions per byte slot however long a segment is, and text at zstd
// level 3 has 8-byte segments.  Here one LANE executes one sequence: its literal run (<= kS2Lit bytes) byte by byte from the
// block's literals, its match (<= kS2Ml bytes, source entirely below the round: two aligned 16-byte loads into the lane's
// scratch, then byte moves inside shared memory).  Sequences that do not fit that -- a source inside the round's own output
// (it is not in memory yet), longer runs -- are left to the whole warp afterwards, one at a time in sequence order, reading
// staged bytes or memory byte by byte.  The staged bytes then leave as 16-byte units, 512 contiguous bytes per warp store.
// Returns false (nothing touched) when too many sequences would be left to the warp.
__device__ __forceinline__ bool x2_round_staged(X2Stage &S, uint8_t *qb, const uint8_t *lit_at, uint32_t qpos, uint32_t round_tot,
                                                bool act, uint32_t ll, uint32_t ml, uint32_t off, uint32_t excl_tot, uint32_t incl_tot,
                                                uint32_t excl_ll, uint32_t lane) {
    const bool lit_small = ll <= kS2Lit;
    const bool m_self = off >= incl_tot && ml <= kS2Ml;  // the source ends at or below the round's first byte, and is short
    const uint32_t deferred = __ballot_sync(kFull, act && !(lit_small && m_self));
    if ((uint32_t)__popc(deferred) > kS2MaxDeferred) return false;
    const uint32_t base16 = qpos & ~15u;
    const uint32_t s_lit = qpos + excl_tot - base16;  // staging index of my literal run; my match follows it
    // --- all loads of the round leave together: per lane the two aligned 16-byte chunks that hold its literal run and the two
    // that hold its match's source; they are parked in the lane's scratch and moved to their place byte by byte in shared
    // memory (a load per byte would expose the memory latency once per byte) ---
    const uint32_t n_lit = (act && lit_small) ? ll : 0u;
    const bool m1 = act && m_self;
    const uint32_t n_m = m1 ? ml : 0u;
    const uint8_t *lsrc = lit_at + excl_ll;
    const uint32_t lo_ = (uint32_t)(reinterpret_cast<uintptr_t>(lsrc) & 15);
    const uint8_t *msrc = qb + (qpos + excl_tot + ll - off);
    const uint32_t mo_ = (uint32_t)(reinterpret_cast<uintptr_t>(msrc) & 15);
    uint4 l0 = make_uint4(0, 0, 0, 0), l1 = l0, m0 = l0, m1v = l0;
    if (n_lit) {
        const uint4 *c = reinterpret_cast<const uint4 *>(lsrc - lo_);
        l0 = c[0];
        if (lo_ + n_lit > 16) l1 = c[1];
    }
    if (m1) {
        const uint4 *c = reinterpret_cast<const uint4 *>(msrc - mo_);
        m0 = c[0];
        if (mo_ + n_m > 16) m1v = c[1];
    }
    uint8_t *my = S.scr + lane * kS2ScrStride;
    *reinterpret_cast<uint4 *>(my) = l0;
    *reinterpret_cast<uint4 *>(my + 16) = l1;
    *reinterpret_cast<uint4 *>(my + 32) = m0;
    *reinterpret_cast<uint4 *>(my + 48) = m1v;
    {
        const uint32_t lmax = __reduce_max_sync(kFull, n_lit);
        const uint8_t *from = my + lo_;
        uint8_t *sp = S.stg + s_lit;
        for (uint32_t k = 0; k < lmax; k++)
            if (k < n_lit) sp[k] = from[k];
        const uint32_t mmax = __reduce_max_sync(kFull, n_m);
        const uint8_t *mfrom = my + 32 + mo_;
        uint8_t *dp = S.stg + s_lit + ll;
        for (uint32_t k = 0; k < mmax; k++)
            if (k < n_m) dp[k] = mfrom[k];
    }
    __syncwarp();
    // --- what is left, the whole warp on one sequence at a time, in order ---
    for (uint32_t D = deferred; D; D &= D - 1) {
        const uint32_t j = (uint32_t)__ffs(D) - 1;
        const uint32_t LL = __shfl_sync(kFull, ll, j), ML = __shfl_sync(kFull, ml, j), OFF = __shfl_sync(kFull, off, j);
        const uint32_t SL = __shfl_sync(kFull, s_lit, j), EL = __shfl_sync(kFull, excl_ll, j);
        const bool done = __shfl_sync(kFull, (uint32_t)m_self, j) != 0;  // its match was short and independent: done above
        if (LL > kS2Lit)
            for (uint32_t k = lane; k < LL; k += 32) S.stg[SL + k] = lit_at[EL + k];
        if (!done) {
            const uint32_t sm0 = SL + LL;         // staging index of the match's first byte
            const uint32_t qm = base16 + sm0;     // its q position
            __syncwarp();                         // earlier sequences' bytes are staged
            if (OFF >= ML || OFF >= 32) {
                for (uint32_t k0 = 0; k0 < ML; k0 += 32) {
                    const uint32_t k = k0 + lane;
                    if (k < ML) {
                        const uint32_t sq = qm + k - OFF;
                        S.stg[sm0 + k] = sq >= qpos ? S.stg[sq - base16] : qb[sq];
                    }
                    if (OFF < ML) __syncwarp();   // the next 32 bytes may repeat these
                }
            } else {  // overlapping with a period below 32: byte k repeats byte k % OFF of the OFF bytes in front of the match
                for (uint32_t k = lane; k < ML; k += 32) {
                    const uint32_t sq = qm - OFF + k % OFF;
                    S.stg[sm0 + k] = sq >= qpos ? S.stg[sq - base16] : qb[sq];
                }
            }
        }
        __syncwarp();
    }
    // --- out: [qpos, qpos + round_tot) as 16-byte units, single bytes at both ends ---
    {
        const uint32_t hi = qpos + round_tot;
        const uint32_t first_full = (qpos + 15) & ~15u, last_full = hi & ~15u;  // full units: [first_full, last_full)
        if (first_full < last_full) {
            for (uint32_t q = first_full + 16 * lane; q < last_full; q += 512)
                *reinterpret_cast<uint4 *>(qb + q) = *reinterpret_cast<const uint4 *>(S.stg + (q - base16));
            const uint32_t head = first_full - qpos;  // < 16
            if (lane < head) qb[qpos + lane] = S.stg[qpos - base16 + lane];
            const uint32_t tail = hi - last_full;  // < 16
            if (lane >= 16 && lane - 16 < tail) qb[last_full + lane - 16] = S.stg[last_full - base16 + lane - 16];
        } else {  // no full unit: at most 30 bytes
            if (lane < round_tot) qb[qpos + lane] = S.stg[qpos - base16 + lane];
        }
    }
    __syncwarp();
    return true;
}

// One frame (status OK, taken by x2_takes), one warp: blocks in order, 32 sequences per round (sequence_execution.go:14-63).
// the consumer's state at the frame's first byte
__device__ __forceinline__ void x2_frame_start(const DeviceBatch &a, uint32_t f, X2State &st) {
    const szb_frame_desc fr = a.frames[f];
    uint8_t *first = a.dst + (fr.nblocks ? a.out_off[fr.first_block] : 0);
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 127);
    st.qb = first - mis;
    st.lit = nullptr;
    st.seen = 0;
    x2_seek(st, mis);
}

// Where the producer's segments go, and who makes the lines.
// X2Inline: the producing warp is the consumer too; it stays one append behind, so that the prefetches it issued for the
// match sources have time to land.
struct X2Inline {
    static constexpr bool kInline = true;
    X2Smem &sm;
    X2State st;
    uint32_t lane, le_mask;
    __device__ __forceinline__ uint32_t line() const { return st.line; }
    __device__ __forceinline__ bool idle(uint32_t prod) const { return prod == st.line + st.head; }  // nothing left in the ring
    __device__ __forceinline__ void set_lit(const uint8_t *lit) { st.lit = lit; }
    __device__ __forceinline__ void seek(uint32_t q) { x2_seek(st, q); }
    __device__ __forceinline__ void flush(uint32_t prod) { st = x2_flush_cold(sm, st, prod, lane, le_mask); }
    __device__ __forceinline__ void drain_to(uint32_t prod) { st = x2_drain_cold(sm, st, prod, lane, le_mask); }
    __device__ __forceinline__ void appended(uint32_t prev_prod, uint32_t) { x2_drain(sm, st, prev_prod, lane, le_mask); }
    __device__ __forceinline__ void finish(uint32_t prod) { st = x2_flush_cold(sm, st, prod, lane, le_mask); }
};

// X2Pair: a second warp of the CTA consumes (k_execute_pair2), for the frames that are far longer than the rest and run almost
// alone at the end: the two warps meet at one __syncthreads per command, and a command is executed by the consumer while the
// producer works on the next round.  The producer only ever needs a LOWER bound of the consumer's line.
enum : uint32_t { kX2cNop = 0, kX2cDrain = 1, kX2cFlush = 2, kX2cSeek = 3, kX2cLit = 4, kX2cExit = 5 };
struct X2PairShared {
    unsigned long long arg[2];  // a q position, or the literals' address
    uint32_t cmd[2];
    uint32_t line[2];           // the consumer's st.line after command i, in slot i & 1
};
struct X2Pair {
    static constexpr bool kInline = false;
    X2PairShared &sh;
    uint32_t lane, it;
    uint32_t seen_line;  // what the consumer had reached one command ago
    __device__ __forceinline__ void hand(uint32_t cmd, unsigned long long arg) {
        if (lane == 0) {
            sh.cmd[it & 1] = cmd;
            sh.arg[it & 1] = arg;
        }
        __syncthreads();  // command `it` starts; command it-1 is complete and has published its line before this barrier
        if (it) seen_line = sh.line[(it - 1) & 1];
        it++;
    }
    __device__ __forceinline__ uint32_t line() const { return seen_line; }
    __device__ __forceinline__ bool idle(uint32_t) const { return false; }
    __device__ __forceinline__ void set_lit(const uint8_t *lit) { hand(kX2cLit, reinterpret_cast<uintptr_t>(lit)); }
    __device__ __forceinline__ void seek(uint32_t q) {
        hand(kX2cSeek, q);
        seen_line = q & ~127u;
    }
    __device__ __forceinline__ void flush(uint32_t prod) {  // returns when everything below prod is in memory
        hand(kX2cFlush, prod);
        hand(kX2cNop, 0);
    }
    __device__ __forceinline__ void drain_to(uint32_t prod) { hand(kX2cDrain, prod); }
    __device__ __forceinline__ void appended(uint32_t, uint32_t prod) { hand(kX2cDrain, prod); }
    __device__ __forceinline__ void finish(uint32_t prod) {
        hand(kX2cFlush, prod);
        hand(kX2cExit, 0);
    }
};

// kDict: the batch is decoded with a dictionary (its own instantiation: the plain one pays nothing for it).
template <bool kDict, class Sink>
__device__ __forceinline__ void x2_frame(const DeviceBatch &a, uint32_t f, X2Smem &sm, X2Stage &stage, Sink &sink, uint32_t lane) {
    const szb_frame_desc fr = a.frames[f];
    const uint32_t b0 = fr.first_block, nb = a.frame_nexec ? a.frame_nexec[f] : fr.nblocks;  // k_frame_verdict
    if (nb == 0) return;
    int err = SZB_OK;
    const uint64_t frame_base = a.out_off[b0];
    const uint32_t lt_mask = 0x7FFFFFFFu >> (31 - lane);  // bits 0..lane-1
    uint8_t *const first = a.dst + frame_base;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 127);  // q position of the frame's first byte
    uint8_t *const qb = first - mis;                                             // q position 0 (x2_frame_start gave the sink the same)
    uint32_t prod = mis;  // segments cover the output up to this q position
    uint32_t nseg = 0;    // segments appended so far

    History hist{1, 4, 8};  // framedecompressor.go:48,59
    // With a dictionary (not a reference behaviour; RFC 8878 section 5) the frame starts with the dictionary's repeat offsets
    // and its content is history in front of the frame: a match may reach up to dlen bytes in front of position 0.
    const uint32_t dlen = (kDict && a.frame_dict && a.frame_dict[f]) ? a.dict_len : 0;
    if (kDict && a.frame_dict && a.frame_dict[f]) hist = History{a.dict_rep[0], a.dict_rep[1], a.dict_rep[2]};
    for (uint32_t bi = 0; bi < nb && err == SZB_OK; bi++) {
        const uint32_t b = b0 + bi;
        const szb_block_desc d = a.blocks[b];
        if (d.type != 2 || d.nseq == 0) continue;  // written by k_execute_bodies already
        const uint8_t *payload = a.src + d.src_off;
        uint32_t qpos = mis + (uint32_t)(a.out_off[b] - frame_base);
        // the consumer is brought up to date at every block start: the literals' base address changes, and blocks in between
        // may have been written elsewhere
        sink.flush(prod);
        if (qpos != prod) {
            sink.seek(qpos);
            prod = qpos;
        }
        // Compressed: ExecuteSequences (sequence_execution.go:14-63)
        const bool lit_rle = d.lit_type == 1;
        // RLE literals: every literal byte is payload[lit_hdr_bytes]; runs read it from that byte's row of the fill table
        const uint8_t *lit = lit_rle ? a.bytefill + 256 * (uint32_t)payload[d.lit_hdr_bytes]
                                     : (d.lit_type == 0 ? payload + d.lit_hdr_bytes : a.litbuf + d.lit_buf_off);
        sink.set_lit(lit);
        const uint32_t nseq = d.nseq;
        uint32_t lit_pos = 0;
        const uint32_t *const seqp = a.seq_ll + d.seq_buf_off;
        // the triples are prefetched to L1 two rounds ahead (one line per array and round): lanes 0..2 take one array each
        const uint32_t *const my_seq = seqp + (lane < 3 ? lane : 0) * a.seq_stride;
        if (lane < 3) {
            SZB_PREFETCH_L1(my_seq);
            if (nseq > 32) SZB_PREFETCH_L1(my_seq + 32);
        }
        for (uint32_t base = 0; base < nseq; base += 32) {
            const uint32_t cnt = nseq - base < 32 ? nseq - base : 32;
            const bool act = lane < cnt;
            const uint32_t *const tr = seqp + (base + lane);  // the arrays are padded to whole rounds: no bounds needed
            uint32_t ll = tr[0], ml = tr[a.seq_stride], ofv = tr[2 * a.seq_stride];
            if (!act) {
                ll = 0;
                ml = 0;
                ofv = 4;
            }
            if (!lit_rle && lane == 0) SZB_PREFETCH_L1(lit + lit_pos + 256);
            if (lane < 3 && base + 64 < nseq) SZB_PREFETCH_L1(my_seq + base + 64);

            // --- offsets through the 3-entry history (nextOffset): the walk jumps from repeat code to repeat code, the run of
            // direct sequences in between is folded in at once (its last three offsets are the new history) ---
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
            const uint32_t excl_tot = incl_tot - tot;  // my literal run starts at qpos + excl_tot
            const uint32_t excl_ll = incl_ll - ll;     // and reads the literals from lit_pos + excl_ll
            // (positions stay below the frame's length, which is the sum stage 3 made of the same numbers: once the literals are
            // known not to run dry, literals + matches of the block so far <= lit_regen + the block's match bytes = out_size)
            {
                // every match must lie inside the frame (ringbuffer.go:203-214); a match length of 0 cannot come out of
                // stage 3 (ML codes start at 3, predefined.go:36-50) and would break the segment count below
                const uint32_t fb = qpos - mis;  // frame bytes in front of the round
                const bool bad = off > fb + excl_tot + ll + dlen;
                if (__any_sync(kFull, act && (bad || off == 0 || ml == 0))) {
                    err = SZB_ERR_CANT_REPEAT_BYTES;
                    break;
                }
            }
            // a match that starts in the dictionary's content is copied by the whole warp, one sequence at a time (below)
            const bool reach = kDict && dlen != 0 && act && off > (qpos - mis) + excl_tot + ll;
            const bool any_reach = kDict && dlen != 0 && __any_sync(kFull, reach);

            // --- rounds of short sequences are staged in shared memory, one lane per sequence (x2_round_staged) ---
            if (SZB_X2_STAGED && Sink::kInline && !lit_rle && !any_reach && round_tot <= kS2Cap) {
                if (!sink.idle(prod)) sink.flush(prod);  // segments still in the ring
                if (x2_round_staged(stage, qb, lit + lit_pos, qpos, round_tot, act, ll, ml, off, excl_tot, incl_tot, excl_ll, lane)) {
                    qpos += round_tot;
                    lit_pos += round_ll;
                    prod = qpos;
                    sink.seek(prod);
                    continue;
                }
            }
            // --- the round's segments go to the ring, as many sequences at a time as the bitmap holds (normally all) ---
            uint32_t start = 0;
            while (start < cnt) {
                uint32_t out_rel = qpos - sink.line();
                if (out_rel + round_tot > kX2Span && prod - sink.line() >= 128) {
                    // rounds of long sequences: the consumer catches up before the round is cut into pieces
                    if (Sink::kInline) sink.drain_to(prod); else sink.flush(prod);
                    out_rel = qpos - sink.line();
                }
                const uint32_t my_rel = out_rel + excl_tot;  // my literal run, relative to the line being consumed
                uint32_t nfit;
                if (start == 0 && !lit_rle && !any_reach && out_rel + round_tot <= kX2Span) {
                    nfit = cnt;  // the usual case: the whole round fits the ring
                } else {
                    const bool fits = my_rel + tot <= kX2Span && !(lit_rle && ll > kX2ConstRun) && !reach;
                    const uint32_t fitmask = __ballot_sync(kFull, fits && lane < cnt) >> start;
                    nfit = fitmask == (0xFFFFFFFFu >> start) ? 32 - start : __ffs(~fitmask) - 1;  // leading fits
                }
                if (nfit == 0) {
                    // one sequence longer than the ring: the whole warp on its literals, then on its match
                    sink.flush(prod);
                    const uint32_t L = __shfl_sync(kFull, ll, start), ML = __shfl_sync(kFull, ml, start);
                    const uint32_t OFF = __shfl_sync(kFull, off, start);
                    const uint32_t Dq = qpos + __shfl_sync(kFull, excl_tot, start);
                    uint8_t *D = qb + Dq;
                    if (lit_rle)
                        warp_memset(D, lit[0], L, lane);
                    else
                        warp_memcpy(D, lit + lit_pos + __shfl_sync(kFull, excl_ll, start), L, lane);
                    __syncwarp();
                    uint8_t *MD = D + L;
                    const uint8_t *MS = MD - OFF;
                    const uint32_t mpos = Dq + L - mis;  // the match's first byte, counted from the frame's
                    if (kDict && OFF > mpos) {
                        // The match starts in the dictionary: byte k repeats history byte (mpos + k - OFF); from k = OFF on that
                        // is a byte of this match, i.e. byte k % OFF of its first OFF bytes.  All sources lie in front of the
                        // match (dictionary content or output already in memory): no order among the lanes is needed.
                        const uint8_t *dict_end = a.dict_content + dlen;  // "frame position 0" of the dictionary's content
                        const uint8_t *frame0 = qb + mis;
                        for (uint32_t k = lane; k < ML; k += 32) {
                            const int32_t sp = (int32_t)(mpos + (k < OFF ? k : k % OFF)) - (int32_t)OFF;
                            MD[k] = sp < 0 ? dict_end[sp] : frame0[sp];
                        }
                    } else if (OFF >= 32) {
                        for (uint32_t k0 = 0; k0 < ML; k0 += 32) {
                            const uint32_t k = k0 + lane;
                            if (k < ML) MD[k] = MS[k];
                            __syncwarp();
                        }
                    } else if (ML) {  // overlapping: periodic extension of the OFF bytes before the match
                        for (uint32_t k = lane; k < ML; k += 32) MD[k] = MS[k % OFF];
                    }
                    __syncwarp();
                    prod = Dq + L + ML;
                    sink.seek(prod);
                    start++;
                    continue;
                }
                const uint32_t end = start + nfit;
                const bool in = lane >= start && lane < end;
                const uint32_t no_lit = __ballot_sync(kFull, in && ll == 0);
                if (in) {
                    uint32_t ord = nseg + 2 * (lane - start) - __popc(no_lit & lt_mask);
                    const uint32_t q_lit = qpos + excl_tot;  // my literal run
                    if (ll) {
                        // literal byte at q: lit[lit_pos + excl_ll + (q - q_lit)]; a run of RLE literals reads the first bytes of the fill row
                        sm.seg[ord & (kX2Ring - 1)] = kX2Lit | (q_lit - (lit_rle ? 0u : lit_pos + excl_ll));
                        atomicOr(&sm.bits[(q_lit >> 5) & (kX2Bits / 32 - 1)], 1u << (q_lit & 31));
                        ord++;
                    }
                    {
                        const uint32_t q_m = q_lit + ll;
                        sm.seg[ord & (kX2Ring - 1)] = off;
                        atomicOr(&sm.bits[(q_m >> 5) & (kX2Bits / 32 - 1)], 1u << (q_m & 31));
                        // the consumer gets here about a round later: have the source on its way to L1
                        SZB_PREFETCH_L1(qb + (q_m - off));
                    }
                }
                nseg += 2 * nfit - __popc(no_lit);
                const uint32_t prev_prod = prod;
                prod = qpos + __shfl_sync(kFull, excl_tot + tot, end - 1);
                __syncwarp();
                sink.appended(prev_prod, prod);  // inline: one append behind, so that the prefetches have time to land
                start = end;
            }
            qpos += round_tot;
            lit_pos += round_ll;
        }
        if (err != SZB_OK) break;
        // trailing literals (sequence_execution.go:55-60, literals.go:411-420): one more segment, or a bulk copy
        const uint32_t rest = d.lit_regen - lit_pos;
        if (rest && (prod - sink.line()) + rest <= kX2Span && !(lit_rle && rest > kX2ConstRun)) {
            if (lane == 0) {
                sm.seg[nseg & (kX2Ring - 1)] = kX2Lit | (qpos - (lit_rle ? 0u : lit_pos));
                atomicOr(&sm.bits[(prod >> 5) & (kX2Bits / 32 - 1)], 1u << (prod & 31));
            }
            nseg++;
            prod += rest;
            __syncwarp();
            sink.drain_to(prod);
        } else if (rest) {
            sink.flush(prod);
            if (lit_rle)
                warp_memset(qb + qpos, lit[0], rest, lane);
            else
                warp_memcpy(qb + qpos, lit + lit_pos, rest, lane);
            __syncwarp();
            prod = qpos + rest;
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

// One warp per frame.  Frames exec_list[first_slot, first_slot + n_slots) that x2_takes; k_execute takes the others.
// ONE warp per CTA (r02i, 65 536 text frames: 9.0 ms against 10.5 ms with four): a CTA's registers and shared memory come back
// when its frame ends instead of when the longest of four frames ends, and the warp's shared memory sits at a constant address.
// 32 CTAs per SM is what the register file holds at 64 registers anyway.
#ifndef SZB_EXEC2_WARPS
#define SZB_EXEC2_WARPS 1
#endif
#ifndef SZB_EXEC2_MIN_CTAS
#define SZB_EXEC2_MIN_CTAS 32
#endif
constexpr int kX2Warps = SZB_EXEC2_WARPS;
template <bool kDict>
__global__ void __launch_bounds__(kX2Warps * 32, SZB_EXEC2_MIN_CTAS) k_execute2(DeviceBatch a, uint32_t first_slot, uint32_t n_slots) {
    __shared__ X2Smem smem[kX2Warps];
    __shared__ X2Stage stages[SZB_X2_STAGED ? kX2Warps : 1];  // (one unused slot when the staged path is compiled out)
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t slot = blockIdx.x * kX2Warps + (threadIdx.x >> 5);
    if (slot >= n_slots) return;
    const uint32_t f = a.exec_list[first_slot + slot];
    if (a.frame_nexec ? a.frame_nexec[f] == 0 : a.frame_status[f] != SZB_OK) return;  // k_frame_verdict
    if (!x2_takes(a, f)) return;              // k_execute's
    X2Smem &sm = smem[threadIdx.x >> 5];
    for (uint32_t wd = lane; wd < kX2Bits / 32; wd += 32) sm.bits[wd] = 0;
    __syncwarp();
    X2Inline sink{sm, X2State{}, lane, 0xFFFFFFFFu >> (31 - lane)};
    x2_frame_start(a, f, sink.st);
    x2_frame<kDict>(a, f, sm, stages[SZB_X2_STAGED ? (threadIdx.x >> 5) : 0], sink, lane);
}

// The frames with the most sequences (exec_list[first_slot, ...): >= 65 536 sequences, those the block-parallel path does not
// take) finish last and then run almost alone: two warps per frame, one producing segments, one making the lines.  Stage 4 of
// the mixed corpus is the time ONE such frame takes.  Same producer, same step as k_execute2.
__global__ void __launch_bounds__(64) k_execute_pair2(DeviceBatch a, uint32_t first_slot, uint32_t n_slots) {
    __shared__ X2Smem sm;
    __shared__ X2Stage stage_unused[1];
    __shared__ X2PairShared sh;
    const uint32_t lane = threadIdx.x & 31;
    if (blockIdx.x >= n_slots) return;
    const uint32_t f = a.exec_list[first_slot + blockIdx.x];
    if (a.frame_status[f] != SZB_OK) return;                // k_frame_verdict; both warps agree
    if (long_jump_ok(a, first_slot + blockIdx.x)) return;   // taken by the block-parallel path (execute_long.cuh)
    if (!(a.pair2 == 1 && x2_takes(a, f))) return;          // 2 GiB and more: k_execute_pair's; pair2 == 2: k_execute_team's
    if ((a.frame_nexec ? a.frame_nexec[f] : a.frames[f].nblocks) == 0) return;  // x2_frame would leave before its first command
    for (uint32_t wd = threadIdx.x; wd < kX2Bits / 32; wd += 64) sm.bits[wd] = 0;
    __syncthreads();
    X2State st0;
    x2_frame_start(a, f, st0);
    if (threadIdx.x < 32) {
        X2Pair sink{sh, lane, 0, st0.line};
        x2_frame<false>(a, f, sm, stage_unused[0], sink, lane);
    } else {
        const uint32_t le_mask = 0xFFFFFFFFu >> (31 - lane);
        X2State st = st0;
        for (uint32_t it = 0;; it++) {
            __syncthreads();
            const uint32_t cmd = sh.cmd[it & 1];
            const unsigned long long arg = sh.arg[it & 1];
            if (cmd == kX2cExit) break;
            if (cmd == kX2cDrain)
                x2_drain(sm, st, (uint32_t)arg, lane, le_mask);
            else if (cmd == kX2cFlush)
                x2_flush(sm, st, (uint32_t)arg, lane, le_mask);
            else if (cmd == kX2cSeek)
                x2_seek(st, (uint32_t)arg);
            else if (cmd == kX2cLit)
                st.lit = reinterpret_cast<const uint8_t *>((uintptr_t)arg);
            if (lane == 0) sh.line[it & 1] = st.line;
        }
    }
}

// ---- k_execute_team: the consumer on several warps -------------------------------------------------------------------------
// A long frame waits for ONE in-order instruction stream: its consumer (k_execute_pair2: ~190 instructions per 128-byte line on a
// warp that runs almost alone, ~9 cycles each).  Here kTeam consumer warps make every line together, each the 32-byte chunks
// w, w + kTeam, ... of it: the gather of a line is 4 / kTeam loads deep instead of four.  Bytes that repeat bytes of the same step
// are exchanged through shared memory: every warp publishes the bytes it has, the others take them from there, chains of
// in-line sources by pointer jumping over the published source positions.  The warps meet at a named barrier (bar.sync 1):
// once per line for the vote "does any byte repeat a byte of this step", twice per exchange round, once after the stores (the
// next line's loads may want them).  The producer (warp 0, x2_frame<X2Pair>) is not part of those; it hands its commands over
// at __syncthreads like k_execute_pair2's.
// MEASURED SLOWER (profiles/r03g_*, SZB_PAIR2=2; off by default): stage 4 of the mixed corpus 18.7 ms with two consumer warps,
// 25.1 ms with four, 20.6 ms with one, against k_execute_pair2's 14.4 ms; one 64 MiB frame 272 ms against 244 ms.  A line's time
// is not the gather's instruction count: it is one dependent trip bitmap -> entry -> source byte -> store per line, which the
// team pays just the same, plus its barriers (the one-warp team, the same work as k_execute_pair2 with the exchange through
// shared memory instead of shuffles, is 43 % slower).
#ifndef SZB_X2_TEAM
#define SZB_X2_TEAM 2
#endif
constexpr uint32_t kX2Team = SZB_X2_TEAM;          // consumer warps per frame: 1, 2 or 4
constexpr uint32_t kX2TeamCpw = 4 / kX2Team;       // 32-byte chunks of a line per consumer warp
static_assert(kX2Team == 1 || kX2Team == 2 || kX2Team == 4, "a line has four chunks");

#if defined(SZB_WARPSIM)
#define SZB_TEAM_SYNC() __named_barrier(1, kX2Team * 32)
#else
#define SZB_TEAM_SYNC() asm volatile("bar.sync 1, %0;" ::"n"(kX2Team * 32) : "memory")
#endif

struct X2TeamShared {
    __align__(16) uint8_t lineb[128];  // the line's bytes as far as they are known
    uint8_t srcpos[128];               // per byte that is not known yet: the place in the line of a byte with the same value
    uint32_t have[4];                  // per chunk: lanes whose byte is in lineb
    uint32_t vote[2][4];               // the team's votes (two in flight: a warp may be one vote ahead of another's read)
};

// true when `p` holds on any thread of the team; every warp of the team calls it the same number of times
__device__ __forceinline__ bool x2t_any(X2TeamShared &T, uint32_t &phase, uint32_t w, uint32_t lane, bool p) {
    const uint32_t b = __ballot_sync(kFull, p);
    if (lane == 0) T.vote[phase][w] = b;
    SZB_TEAM_SYNC();
    uint32_t r = 0;
#pragma unroll
    for (uint32_t k = 0; k < kX2Team; k++) r |= T.vote[phase][k];
    phase ^= 1;
    return r != 0;
}

// The bytes [lo, hi) of the line at st.line, my chunks of it (x2_step for a team).
template <bool kWhole>
__device__ __forceinline__ void x2t_step(X2Smem &sm, X2TeamShared &T, X2State &st, uint32_t &phase, uint32_t lo, uint32_t hi, uint32_t w,
                                         uint32_t lane, uint32_t le_mask) {
    uint32_t *bw = &sm.bits[(st.line >> 5) & (kX2Bits / 32 - 1)];
    const uint4 m4 = *reinterpret_cast<const uint4 *>(bw);
    const uint32_t m[4] = {m4.x, m4.y, m4.z, m4.w};
    const uint32_t cnt[4] = {(uint32_t)__popc(m[0]), (uint32_t)__popc(m[1]), (uint32_t)__popc(m[2]), (uint32_t)__popc(m[3])};
    uint32_t v[kX2TeamCpw], so[kX2TeamCpw];
    bool ing[kX2TeamCpw], live[kX2TeamCpw];
    bool any_ing = false;
#pragma unroll
    for (uint32_t k = 0; k < kX2TeamCpw; k++) {
        const uint32_t c = w + k * kX2Team;  // my k-th chunk of the line
        const uint32_t rel = (c << 5) + lane;
        uint32_t before = 0;  // segments that start in the chunks in front of this one
#pragma unroll
        for (uint32_t j = 0; j < 3; j++) before += j < c ? cnt[j] : 0u;
        const uint32_t mc = c == 0 ? m[0] : (c == 1 ? m[1] : (c == 2 ? m[2] : m[3]));  // (no indexed register arrays)
        const uint32_t ord = st.seen - 1 + before + __popc(mc & le_mask);  // the last segment that starts at or before my byte
        const uint32_t e = sm.seg[ord & (kX2Ring - 1)];
        live[k] = kWhole || (rel >= lo && rel < hi);
        const uint32_t idx = st.line + rel - (e & ~kX2Lit);
        ing[k] = live[k] && e <= rel - lo;  // a match whose source is a byte of this step: not in memory yet
        so[k] = rel - e;
        v[k] = 0;
        if (live[k] && !ing[k]) v[k] = ((int32_t)e < 0 ? st.lit : st.qb)[idx];
        any_ing |= ing[k];
    }
    st.seen += cnt[0] + cnt[1] + cnt[2] + cnt[3];
    if (x2t_any(T, phase, w, lane, any_ing)) {
        // publish what is known; a source position is >= lo and below its reader, i.e. a live byte of this step
#pragma unroll
        for (uint32_t k = 0; k < kX2TeamCpw; k++) {
            const uint32_t c = w + k * kX2Team;
            const uint32_t rel = (c << 5) + lane;
            T.lineb[rel] = (uint8_t)v[k];
            T.srcpos[rel] = (uint8_t)so[k];
            const uint32_t known = __ballot_sync(kFull, !ing[k]);
            if (lane == 0) T.have[c] = known;
        }
        SZB_TEAM_SYNC();
        for (;;) {
            // read: a known source gives its value; an unknown one gives ITS source (the same value, further down)
            bool got[kX2TeamCpw];
            bool open = false;
#pragma unroll
            for (uint32_t k = 0; k < kX2TeamCpw; k++) {
                got[k] = false;
                if (ing[k]) {
                    if ((T.have[so[k] >> 5] >> (so[k] & 31)) & 1) {
                        v[k] = T.lineb[so[k]];
                        ing[k] = false;
                        got[k] = true;
                    } else {
                        so[k] = T.srcpos[so[k]];
                    }
                }
                open |= ing[k];
            }
            const bool more = x2t_any(T, phase, w, lane, open);  // its barrier: every read of the round is done
#pragma unroll
            for (uint32_t k = 0; k < kX2TeamCpw; k++) {
                const uint32_t c = w + k * kX2Team;
                const uint32_t rel = (c << 5) + lane;
                if (got[k]) T.lineb[rel] = (uint8_t)v[k];
                if (ing[k]) T.srcpos[rel] = (uint8_t)so[k];
                const uint32_t newly = __ballot_sync(kFull, got[k]);
                if (lane == 0 && newly) T.have[c] |= newly;
            }
            if (!more) break;
            SZB_TEAM_SYNC();  // the round's writes are in place
        }
    }
    uint8_t *out = st.qb + st.line + lane;
#pragma unroll
    for (uint32_t k = 0; k < kX2TeamCpw; k++)
        if (live[k]) out[(w + k * kX2Team) << 5] = (uint8_t)v[k];
    if (w == 0 && lane < 4) bw[lane] = 0;  // every warp read the words before the vote's barrier; the bitmap is a ring
    SZB_TEAM_SYNC();  // the stores: the next step's loads (another warp's) may want these bytes
}

__device__ __forceinline__ void x2t_drain(X2Smem &sm, X2TeamShared &T, X2State &st, uint32_t &phase, uint32_t limit, uint32_t w, uint32_t lane,
                                          uint32_t le_mask) {
    uint32_t n = (limit - st.line) >> 7;
    if (n == 0) return;
    if (st.head) {  // the rest of a line that was flushed in part
        x2t_step<false>(sm, T, st, phase, st.head, 128, w, lane, le_mask);
        st.line += 128;
        st.head = 0;
        n--;
    }
    for (; n; n--) {
        x2t_step<true>(sm, T, st, phase, 0, 128, w, lane, le_mask);
        st.line += 128;
    }
}
__device__ __forceinline__ void x2t_flush(X2Smem &sm, X2TeamShared &T, X2State &st, uint32_t &phase, uint32_t prod, uint32_t w, uint32_t lane,
                                          uint32_t le_mask) {
    x2t_drain(sm, T, st, phase, prod, w, lane, le_mask);
    const uint32_t hi = prod - st.line;  // < 128
    if (hi > st.head) {
        x2t_step<false>(sm, T, st, phase, st.head, hi, w, lane, le_mask);
        st.head = hi;
    }
}

#ifndef SZB_TEAM_MIN_CTAS
#define SZB_TEAM_MIN_CTAS 11
#endif
__global__ void __launch_bounds__((kX2Team + 1) * 32, SZB_TEAM_MIN_CTAS) k_execute_team(DeviceBatch a, uint32_t first_slot, uint32_t n_slots) {
    __shared__ X2Smem sm;
    __shared__ X2Stage stage_unused[1];
    __shared__ X2PairShared sh;
    __shared__ X2TeamShared T;
    const uint32_t lane = threadIdx.x & 31;
    if (blockIdx.x >= n_slots) return;
    const uint32_t f = a.exec_list[first_slot + blockIdx.x];
    if (a.frame_status[f] != SZB_OK) return;                // k_frame_verdict; every warp agrees
    if (long_jump_ok(a, first_slot + blockIdx.x)) return;   // taken by the block-parallel path (execute_long.cuh)
    if (!(a.pair2 == 2 && x2_takes(a, f))) return;          // 2 GiB and more: k_execute_pair's
    if ((a.frame_nexec ? a.frame_nexec[f] : a.frames[f].nblocks) == 0) return;  // x2_frame would leave before its first command
    for (uint32_t wd = threadIdx.x; wd < kX2Bits / 32; wd += (kX2Team + 1) * 32) sm.bits[wd] = 0;
    __syncthreads();
    X2State st0;
    x2_frame_start(a, f, st0);
    if (threadIdx.x < 32) {
        X2Pair sink{sh, lane, 0, st0.line};
        x2_frame<false>(a, f, sm, stage_unused[0], sink, lane);
    } else {
        const uint32_t w = (threadIdx.x >> 5) - 1;  // my place in the team
        const uint32_t le_mask = 0xFFFFFFFFu >> (31 - lane);
        X2State st = st0;  // every warp of the team keeps the same state
        uint32_t phase = 0;
        for (uint32_t it = 0;; it++) {
            __syncthreads();
            const uint32_t cmd = sh.cmd[it & 1];
            const unsigned long long arg = sh.arg[it & 1];
            if (cmd == kX2cExit) break;
            if (cmd == kX2cDrain)
                x2t_drain(sm, T, st, phase, (uint32_t)arg, w, lane, le_mask);
            else if (cmd == kX2cFlush)
                x2t_flush(sm, T, st, phase, (uint32_t)arg, w, lane, le_mask);
            else if (cmd == kX2cSeek)
                x2_seek(st, (uint32_t)arg);
            else if (cmd == kX2cLit)
                st.lit = reinterpret_cast<const uint8_t *>((uintptr_t)arg);
            if (w == 0 && lane == 0) sh.line[it & 1] = st.line;
        }
    }
}
