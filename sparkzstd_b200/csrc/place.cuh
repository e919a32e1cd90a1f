// place.cuh -- stage 4 for the frames one warp executes (sm_100a): k_resolve + k_place.
// Included by execute.cuh inside namespace szb.  Plain CUDA C++ (tests/host_sim runs both kernels on an emulated CTA).
//
// replaces: decompression/sequence_execution.go:14-114 (ExecuteSequences, nextOffset) and ringbuffer.go:197-277 (match copy)
//
// What is sequential about executing a frame -- the repeat-offset history, the running output and literal positions, "the
// first error wins" -- costs a handful of instructions per sequence when ONE LANE walks the frame's sequences in order, and
// ~230 warp instructions per 32 sequences when a warp does it with scans (k_execute's producer half).  So the work is cut
// differently here:
//
//   k_resolve   one LANE per frame, sequential over its blocks and sequences.  Every sequence becomes up to two SEGMENTS
//               -- its literal run (if any) and its match: a contiguous piece of output with a contiguous source -- and
//               a segment is one 32-bit entry in the block's list plus one bit, at its first output position, in a
//               bitmap over the output.  A match's entry is its offset, through the repeat history; a literal run's is
//               bit 31 | the match bytes of the block in front of it (what a literal byte's position and its index into
//               the block's literals differ by).  The reference's checks are applied in the reference's order (entropy
//               status of the block, literals run dry, match reaching in front of the frame): the first failing step
//               decides.
//   k_place     one WARP per frame, the output produced in address order, one aligned 128-byte line per step, lane i
//               making bytes i, 32+i, 64+i, 96+i.  A byte finds its segment with one popcount over the bitmap word of its
//               32-byte chunk (segments that start at or before it), reads ONE entry and ONE source byte (a literal,
//               or the output `offset` bytes back, which an earlier line, block or kernel has written).  A source inside the line is another byte of the step: it is taken from the lane
//               that holds it with a shuffle, chains of them (overlapping matches) by pointer jumping over shuffles.
//               No shared memory, no atomics, no scans: per line ~100 warp instructions where k_execute needs ~270.
//
// Bit positions: byte p of the output (p counted from DeviceBatch::dst) has "q position" p + (dst & 127), so that a line is
// an aligned 128-byte line of memory; inside a frame positions are taken relative to line0, the line holding the frame's
// first byte, and fit 32 bits.  Frame f's bitmap words start at (line0 >> 5) + 4 f: frames never share a word, and a
// line's four words are one aligned 16-byte load.
#pragma once

constexpr int32_t kPlaceFallback = 1;  // place_state: the frame is k_execute's (more output than the bitmap holds, or >= 2 GiB)
constexpr uint32_t kRecLiteral = 0x80000000u;
#ifndef SZB_PLACE_PREFETCH_AHEAD
#define SZB_PLACE_PREFETCH_AHEAD 64
#endif

// Keeps a base address (or a 64-bit value) whole in its registers: without it the compiler starts every index computation
// again from the kernel's parameters (four instructions per access instead of one).
#if defined(__CUDA_ARCH__)
#define SZB_KEEP(v) asm volatile("" : "+l"(v))
#define SZB_KEEP_PTR(p)               \
    asm volatile("" : "+l"(p));       \
    __builtin_assume(__isGlobal(p))
#else
#define SZB_KEEP(v) ((void)(v))
#define SZB_KEEP_PTR(p) ((void)(p))
#endif

// every kernel evaluates the same predicate
__device__ __forceinline__ bool place_on(const DeviceBatch &a) { return a.rec != nullptr && a.total[0] <= a.bm_bound; }

// k_resolve's two outputs for one frame: the bitmap words and the segment entries.  A lane writing to its own frame's
// memory costs a sector per lane and store, and the walk then waits on its stores; so both outputs are staged in the
// warp's shared memory for a round of 32 sequences and leave together, a row (= a lane's frame) at a time, coalesced.
// The bitmap is zero before k_resolve runs (k_place_zero): only words that hold a bit are stored.
#ifndef SZB_RESOLVE_ROUND
#define SZB_RESOLVE_ROUND 16
#endif
constexpr uint32_t kRound = SZB_RESOLVE_ROUND;  // sequences a lane walks per round (16 or 32): shared memory per warp is what decides
                                                // how many warps an SM holds, and every frame wants its lane at once
static_assert(kRound == 16 || kRound == 32, "a round is 16 or 32 sequences");
constexpr uint32_t kEntRow = 2 * kRound + 1;  // two slots per sequence of a round + the literals after a block's last sequence; 0 = no entry
constexpr uint32_t kBitWords = kRound / 2;    // completed bitmap words of a round that are staged (32 sequences of text cover ~11); more
                                              // go straight to memory
constexpr uint32_t kBitRow = kBitWords + 1;
struct SegWriter {
    uint32_t *bm;      // the frame's bitmap words
    uint32_t widx;     // word being assembled
    uint32_t w;
    uint32_t wfirst;   // widx when the round began: staged word k is word wfirst + k
    uint32_t wmask;    // staged words of this round
    uint32_t *brow;    // my row of staged words
    uint32_t *ent;     // my row of staged entries
    uint32_t n;        // entries of the block so far
};
// a segment starts at q position x (>= every earlier one of the frame); no branch
__device__ __forceinline__ void sw_mark(SegWriter &s, uint32_t x, bool on) {
    const uint32_t i = x >> 5;
    const bool flush = on && i != s.widx;
    const uint32_t slot = s.widx - s.wfirst;
    if (flush && slot < kBitWords) s.brow[slot] = s.w;
    if (flush && slot >= kBitWords) s.bm[s.widx] = s.w;
    if (flush && slot < kBitWords) s.wmask |= 1u << slot;
    if (flush) s.w = 0;
    if (on) s.widx = i;
    if (on) s.w |= 1u << (x & 31);
}

// The sequential state of a frame while its sequences are walked.
struct ResolveState {
    uint32_t h0, h1, h2;  // repeat-offset history (sequence_execution.go:65-114)
    uint32_t L, CM;       // literal bytes / match bytes of the block so far
};

// One sequence, no branches: its offset through the history (nextOffset, sequence_execution.go:65-114), its two segments.
// Returns non-zero when the reference would fail here (the caller then walks the block again with resolve_exact).
__device__ __forceinline__ uint32_t resolve_one(ResolveState &r, SegWriter &w, uint32_t k, uint32_t ll, uint32_t ml, uint32_t ofv,
                                                uint32_t regen, uint32_t fb, uint32_t bsq, uint32_t last_q) {
    const bool rep = ofv <= 3;
    const uint32_t idx = ofv - 1 + (ll == 0 ? 1u : 0u);  // 0: h0, 1: h1, 2: h2, 3: h0 - 1
    const uint32_t pick = idx == 0 ? r.h0 : (idx == 1 ? r.h1 : (idx == 2 ? r.h2 : r.h0 - 1));
    const uint32_t off = rep ? pick : ofv - 3;
    const bool shift = !rep || idx >= 2;  // the oldest entry leaves
    const bool keep = rep && idx == 0;    // nothing moves
    r.h2 = shift ? r.h1 : r.h2;
    r.h1 = keep ? r.h1 : r.h0;
    r.h0 = off;
    // literals run dry (sequence_execution.go:19-34); the match must lie inside the frame (ringbuffer.go:203-214)
    const uint32_t Ln = r.L + ll;
    const uint32_t bad = (Ln > regen ? 1u : 0u) | (off == 0 ? 1u : 0u) | (off > fb + Ln + r.CM ? 1u : 0u) | (ml == 0 ? 1u : 0u);
    // positions stay inside the block whatever the sequences say (a failing block is walked again and none of this is used)
    uint32_t ps = bsq + r.L + r.CM, ms = ps + ll;
    ps = ps < last_q ? ps : last_q;
    ms = ms < last_q ? ms : last_q;
    sw_mark(w, ps, ll != 0);
    sw_mark(w, ms, true);
    w.ent[2 * k] = ll ? (r.CM | kRecLiteral) : 0u;
    w.ent[2 * k + 1] = off;
    w.n += ll ? 2u : 1u;
    r.L = Ln;
    r.CM += ml;
    return bad;
}

// The reference's order of checks, one sequence after the other, for a block resolve_one found wrong: which error it is.
__device__ __noinline__ int resolve_exact(const uint32_t *pl, const uint32_t *pm, const uint32_t *po, uint32_t nseq, uint32_t regen,
                                          bool lit_rle, uint32_t fb, uint32_t h0, uint32_t h1, uint32_t h2) {
    uint32_t L = 0, CM = 0;
    bool overrun = false;  // RLE literals never run dry (literals.go:390-396): the reference goes on and panics at the block's end
    for (uint32_t i = 0; i < nseq; i++) {
        const uint32_t ll = pl[i], ml = pm[i], ofv = po[i];
        if (ll && !overrun) {  // sequence_execution.go:19-34, literals.go:383-409
            if (lit_rle) {
                if (L + ll > regen) overrun = true;
            } else if (L == regen) {
                return SZB_ERR_UNEXPECTED_EOF;  // literals.go:398-400 io.EOF
            } else if (regen - L < ll) {
                return SZB_ERR_DIDNT_COPY_ALL_LITERAL_BYTES;  // sequence_execution.go:26-28
            }
        }
        uint32_t off;
        if (ofv > 3) {
            off = ofv - 3;
            h2 = h1;
            h1 = h0;
            h0 = off;
        } else {
            const uint32_t idx = ofv - 1 + (ll == 0 ? 1u : 0u);
            if (idx == 0) {
                off = h0;
            } else if (idx == 1) {
                off = h1;
                h1 = h0;
                h0 = off;
            } else {
                off = idx == 2 ? h2 : h0 - 1;
                h2 = h1;
                h1 = h0;
                h0 = off;
            }
        }
        if (off == 0 || off > (uint64_t)fb + L + ll + CM || ml == 0) return SZB_ERR_CANT_REPEAT_BYTES;  // ringbuffer.go:203-214
        L += ll;
        CM += ml;
    }
    return overrun ? SZB_ERR_PANIC : SZB_OK;  // GetRest: make([]byte, negative) (literals.go:411-420)
}

// The bitmap words the batch's output covers, zeroed (k_resolve only stores words that hold a bit).
__global__ void __launch_bounds__(256) k_place_zero(DeviceBatch a) {
    if (!place_on(a)) return;
    const uint64_t quads = (((a.total[0] + 256) >> 5) + 4ull * a.nframes + 8) >> 2;
    uint4 *q = reinterpret_cast<uint4 *>(a.bm);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (uint64_t)gridDim.x * blockDim.x) q[i] = make_uint4(0, 0, 0, 0);
}

// One lane per frame, after k_scan_blocks and before k_frame_verdict.  The long frames (exec_list[0, n_long)) have their
// own stage 4 (execute_long.cuh, k_execute_pair), and a frame of many sequences would keep the 31 other lanes of its warp
// waiting: exec_list[n_long, n_noplace) stays with k_execute, one warp per frame.
//
// A lane walks its frame's sequences in order, but a lane reading its own frame's arrays would fetch a sector per lane
// and load (32 sectors per warp instruction, and the walk then waits for memory at every step: 9.6 ms for 4 x 10^8
// sequences).  So the warp loads together: per round, for every lane j in turn, all 32 lanes fetch the next 32 sequences
// of lane j's block -- one 128-byte line per array -- into shared memory (rows padded to 33 words: a lane reading along
// its row and the warp writing across a row both hit 32 different banks), then every lane walks its own row.
constexpr int kResolveWarps = 2;
constexpr uint32_t kResolveRow = kRound + 1;
__device__ __forceinline__ void stage_copy(uint32_t *smem_dst, const uint32_t *src) {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
#else
    *smem_dst = *src;
#endif
}
__device__ __forceinline__ void stage_commit() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void stage_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

struct ResolveLane {      // where a lane is in its frame
    uint32_t b, b_end;    // next block to look at; one past the frame's last block
    uint32_t nseq, i;     // current block: sequences, sequences walked
    uint32_t regen, fb, bsq, last_q;
    uint64_t sbo;
    uint32_t bad;
    bool lit_rle;
    bool active;
    int err;
    ResolveState at_start;
    uint32_t *rec;        // the block's entries in memory
};

// moves on to the frame's next block with sequences (the blocks' own verdicts are looked at in order on the way:
// literals, then sequences, framedecompressor.go:93-126); the frame is done when there is none
__device__ __forceinline__ void resolve_next_block(const DeviceBatch &a, ResolveLane &s, ResolveState &r, SegWriter &w, uint64_t frame_start,
                                                   uint32_t A, uint64_t line0) {
    for (;;) {
        if (s.b >= s.b_end) {
            s.active = false;
            return;
        }
        const uint32_t b = s.b++;
        const int ls = a.lit_status[b], ss = a.seq_status[b], hs = a.blocks[b].hdr_status;
        if ((ls | ss | hs) != 0) {
            s.err = ls ? ls : (ss ? ss : hs);
            s.active = false;
            return;
        }
        const szb_block_desc *d = a.blocks + b;
        const uint32_t nseq = d->nseq;
        if (d->type != 2 || nseq == 0) continue;  // k_execute_bodies
        s.nseq = nseq;
        s.i = 0;
        s.regen = d->lit_regen;
        s.lit_rle = d->lit_type == 1;
        s.sbo = d->seq_buf_off;
        const uint64_t bstart = a.out_off[b];
        s.bsq = (uint32_t)(bstart + A - line0);                // q position of the block's first byte
        s.fb = (uint32_t)(bstart - frame_start);               // frame bytes in front of the block
        s.last_q = s.bsq + (uint32_t)a.out_size[b] - 1;        // the block has sequences: it is not empty
        s.bad = 0;
        s.at_start = r;
        r.L = 0;
        r.CM = 0;
        s.rec = a.rec + a.rec_off[b];
        w.n = 0;
        return;
    }
}

struct ResolveSmem {
    uint32_t in[3][32 * kResolveRow];  // the round's sequences: ll | ml | of, a row per lane
    uint32_t ent[32 * kEntRow];
    uint32_t bits[32 * kBitRow];
};

// the next (up to) kRound sequences of every lane's block: three rows per lane, fetched by the whole warp with asynchronous
// copies (all rows of a round in flight together, no register held for them); 32 / kRound rows per step
__device__ __forceinline__ void resolve_fetch(const DeviceBatch &a, ResolveSmem &sm, const ResolveLane &s, uint32_t lane) {
    const uint32_t *row = s.active ? a.seq_ll + s.sbo + s.i : nullptr;
    const uint64_t stride = a.seq_stride;
    constexpr uint32_t kRowsPerStep = 32 / kRound;
    const uint32_t sub = lane / kRound, col = lane % kRound;
#pragma unroll 4
    for (uint32_t j0 = 0; j0 < 32; j0 += kRowsPerStep) {
        const uint32_t j = j0 + sub;
        const uint32_t *pj = reinterpret_cast<const uint32_t *>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(row), j));
        if (pj == nullptr) continue;
        stage_copy(&sm.in[0][j * kResolveRow + col], pj + col);
        stage_copy(&sm.in[1][j * kResolveRow + col], pj + stride + col);
        stage_copy(&sm.in[2][j * kResolveRow + col], pj + 2 * stride + col);
    }
    stage_commit();
}

__global__ void __launch_bounds__(kResolveWarps * 32) k_resolve(DeviceBatch a) {
    __shared__ ResolveSmem smem_all[kResolveWarps];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ResolveSmem &sm = smem_all[warp];
    const uint32_t slot = (blockIdx.x * kResolveWarps + warp) * 32 + lane;
    const uint32_t lt = (1u << lane) - 1;
    ResolveLane s{};
    ResolveState r{1, 4, 8, 0, 0};  // framedecompressor.go:48,59
    SegWriter w{};
    w.brow = sm.bits + lane * kBitRow;
    w.ent = sm.ent + lane * kEntRow;
    uint32_t f = 0, A = 0;
    uint64_t frame_start = 0, line0 = 0;
    if (slot < a.nframes) {
        f = a.exec_list[slot];
        const uint32_t b0 = a.frames[f].first_block, nb = a.frames[f].nblocks;
        if (slot < a.n_noplace || !place_on(a)) {
            a.place_state[f] = kPlaceFallback;
        } else if (nb == 0) {
            a.place_state[f] = SZB_OK;
        } else {
            frame_start = a.out_off[b0];
            const uint64_t frame_len = a.out_off[b0 + nb - 1] + a.out_size[b0 + nb - 1] - frame_start;
            if (frame_len >= 0x7FFFFF00u) {  // positions inside a frame are 31 bits here
                a.place_state[f] = kPlaceFallback;
            } else {
                A = (uint32_t)(reinterpret_cast<uintptr_t>(a.dst) & 127);
                line0 = (frame_start + A) & ~(uint64_t)127;
                w.bm = a.bm + ((line0 >> 5) + 4ull * f);
                w.widx = (uint32_t)(frame_start + A - line0) >> 5;  // w = 0: storing it changes nothing
                s.b = b0;
                s.b_end = b0 + nb;
                s.active = true;
                resolve_next_block(a, s, r, w, frame_start, A, line0);
                if (!s.active) a.place_state[f] = s.err;  // no block with sequences, or a block that failed before
            }
        }
    }
    resolve_fetch(a, sm, s, lane);
    while (__any_sync(kFull, s.active)) {
        stage_wait();
        __syncwarp();
        // what of my row leaves after this round
        uint32_t out_slots = 0;                 // staged entry slots in use; bit 8: and the slot of a block's last literals
        uint32_t *out_rec = nullptr;            // where the row's first entry goes
        if (s.active) {
            const uint32_t cnt = s.nseq - s.i < kRound ? s.nseq - s.i : kRound;
            const uint32_t *ql = sm.in[0] + lane * kResolveRow, *qm = sm.in[1] + lane * kResolveRow, *qo = sm.in[2] + lane * kResolveRow;
            out_rec = s.rec + w.n;
            out_slots = 2 * cnt;
            w.wfirst = w.widx;
            uint32_t k = 0;
            for (; k + 4 <= cnt; k += 4) {
                const uint32_t l0 = ql[k], l1 = ql[k + 1], l2 = ql[k + 2], l3 = ql[k + 3];
                const uint32_t m0 = qm[k], m1 = qm[k + 1], m2 = qm[k + 2], m3 = qm[k + 3];
                const uint32_t o0 = qo[k], o1 = qo[k + 1], o2 = qo[k + 2], o3 = qo[k + 3];
                s.bad |= resolve_one(r, w, k, l0, m0, o0, s.regen, s.fb, s.bsq, s.last_q);
                s.bad |= resolve_one(r, w, k + 1, l1, m1, o1, s.regen, s.fb, s.bsq, s.last_q);
                s.bad |= resolve_one(r, w, k + 2, l2, m2, o2, s.regen, s.fb, s.bsq, s.last_q);
                s.bad |= resolve_one(r, w, k + 3, l3, m3, o3, s.regen, s.fb, s.bsq, s.last_q);
            }
            for (; k < cnt; k++) s.bad |= resolve_one(r, w, k, ql[k], qm[k], qo[k], s.regen, s.fb, s.bsq, s.last_q);
            s.i += cnt;
            if (s.i == s.nseq) {  // the block is walked
                if (s.bad) {
                    const uint32_t *pl = a.seq_ll + s.sbo;
                    s.err = resolve_exact(pl, pl + a.seq_stride, pl + 2 * a.seq_stride, s.nseq, s.regen, s.lit_rle, s.fb, s.at_start.h0,
                                          s.at_start.h1, s.at_start.h2);
                    if (s.err == SZB_OK) s.err = SZB_ERR_PANIC;  // cannot happen: whatever resolve_one flags is an error
                    s.active = false;
                    out_slots = 0;  // nothing of this frame will be read
                    w.wmask = 0;
                } else {
                    // trailing literals (sequence_execution.go:55-60)
                    w.ent[2 * kRound] = 0;
                    if (r.L < s.regen) {
                        sw_mark(w, s.bsq + r.L + r.CM, true);
                        w.ent[2 * kRound] = r.CM | kRecLiteral;
                    }
                    out_slots |= 0x100;
                    resolve_next_block(a, s, r, w, frame_start, A, line0);
                }
                if (!s.active) {  // the frame is done
                    w.bm[w.widx] = w.w;
                    a.place_state[f] = s.err;
                }
            }
        }
        __syncwarp();
        resolve_fetch(a, sm, s, lane);  // the next round's sequences travel while this round's results leave
        // entries: a row at a time, compacted (a sequence without literals leaves an empty slot) and coalesced
        for (uint32_t j = 0; j < 32; j++) {
            const uint32_t oj = __shfl_sync(kFull, out_slots, j);
            if (oj == 0) continue;
            const uint32_t nj = oj & 0xFF;
            uint32_t *rj = reinterpret_cast<uint32_t *>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(out_rec), j));
            const uint32_t *row = sm.ent + j * kEntRow;
            const uint32_t v0 = lane < nj ? row[lane] : 0u;
            const uint32_t b0 = __ballot_sync(kFull, v0 != 0);
            uint32_t c = __popc(b0);
            if (v0) rj[__popc(b0 & lt)] = v0;
            if (kRound == 32) {
                const uint32_t v1 = lane + 32 < nj ? row[lane + 32] : 0u;
                const uint32_t b1 = __ballot_sync(kFull, v1 != 0);
                if (v1) rj[c + __popc(b1 & lt)] = v1;
                c += __popc(b1);
            }
            if (lane == 0 && (oj & 0x100) && row[2 * kRound]) rj[c] = row[2 * kRound];
        }
        // completed bitmap words
        for (uint32_t j = 0; j < 32; j++) {
            const uint32_t mj = __shfl_sync(kFull, w.wmask, j);
            if (mj == 0) continue;
            uint32_t *bj = reinterpret_cast<uint32_t *>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(w.bm + w.wfirst), j));
            if ((mj >> lane) & 1) bj[lane] = sm.bits[j * kBitRow + lane];  // mj has bits below kBitWords only
        }
        w.wmask = 0;
        __syncwarp();
    }
    stage_wait();
}

__device__ __forceinline__ uint32_t range_mask(uint32_t lo, uint32_t hi) {  // bits [lo, hi), 0 <= lo, hi <= 32
    const uint32_t up = hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
    const uint32_t dn = lo >= 32 ? 0xFFFFFFFFu : ((1u << lo) - 1u);
    return up & ~dn;
}

// byte `i` (0..3) of w
__device__ __forceinline__ uint32_t byte_of(uint32_t w, uint32_t i) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0, 0x4440 + i);
#else
    return (w >> (8 * i)) & 0xFF;
#endif
}

// ---- k_place ----
// In-line sources.  vp: my bytes of the chunks that are final, chunk c in byte c; chunk c is being resolved: vc my byte of it
// (final unless ic), sc the line position of its source (< my own position), oc the lanes that wait.  A source in an
// earlier chunk is final in its lane's vp: one shuffle.  A source in the same chunk may itself be waiting (overlapping
// matches): then my source becomes its source (pointer jumping: chains such as offset 1 halve every round).
#define SZB_RESOLVE(c, vc, ic, sc, oc)                                                             \
    if (oc) {                                                                                      \
        uint32_t s = sc;                                                                           \
        bool un = ic;                                                                              \
        uint32_t open = __ballot_sync(kFull, un && (s >> 5) == c && ((oc >> (s & 31)) & 1));       \
        if (!open) {                                                                               \
            const uint32_t w = __shfl_sync(kFull, vp | (vc << (8 * c)), s);                        \
            if (un) vc = byte_of(w, s >> 5);                                                       \
        } else {                                                                                   \
            open = oc;                                                                             \
            do {                                                                                   \
                const uint32_t w = __shfl_sync(kFull, vp | (vc << (8 * c)), s);                    \
                const uint32_t ps = __shfl_sync(kFull, s, s);                                      \
                if (un) {                                                                          \
                    if ((s >> 5) == c && ((open >> (s & 31)) & 1)) {                               \
                        s = ps;                                                                    \
                    } else {                                                                       \
                        vc = byte_of(w, s >> 5);                                                   \
                        un = false;                                                                \
                    }                                                                              \
                }                                                                                  \
                open = __ballot_sync(kFull, un);                                                   \
            } while (open);                                                                        \
        }                                                                                          \
    }                                                                                              \
    vp |= vc << (8 * c);

// One chunk (32 bytes, lane = byte) of a WHOLE line: finds the byte's segment, reads its entry, issues its load.
//   u     the chunk's bitmap word;  nm1: segments of the block that start below the chunk, minus 1
//   pd    my byte in the output; DL: what a literal byte's address differs by from pd - entry
// Returns the byte (when its source is in memory); e: the entry (e <= rel: the source is a byte of this step, rel - e).
template <bool kRle>
__device__ __forceinline__ uint32_t place_whole_chunk(const uint32_t *__restrict__ rec, const uint8_t *pd, long long DL, uint32_t rle_byte,
                                                      uint32_t u, uint32_t nm1, uint32_t rel, uint32_t le, uint32_t &e) {
    e = rec[nm1 + __popc(u & le)];  // the last segment that starts at or before my byte
    const bool is_l = e >= kRecLiteral;
    const uint8_t *p = pd - e;
    if (is_l) p += DL;
    uint32_t v = ld_u8_if(p, (kRle ? !is_l : true) && e > rel);  // e > rel: true for every literal entry
    if (kRle && is_l) v = rle_byte;
    return v;
}

template <bool kRle>
__device__ __forceinline__ void place_whole_line(const uint32_t *__restrict__ rec, uint8_t *pd, long long DL, uint32_t rle_byte, uint4 U,
                                                 uint32_t &nm1, uint32_t lane, uint32_t le) {
    uint32_t e0, e1, e2, e3;
#ifndef SZB_PLACE_NO_PREFETCH
    // the entries are read in order, ~19 per line: the line of memory that holds those of the line after the next is asked for now
    if (lane == 0) SZB_PREFETCH_L1(rec + nm1 + SZB_PLACE_PREFETCH_AHEAD);
#endif
    uint32_t v0 = place_whole_chunk<kRle>(rec, pd, DL, rle_byte, U.x, nm1, lane, le, e0);
    nm1 += __popc(U.x);
    uint32_t v1 = place_whole_chunk<kRle>(rec, pd + 32, DL, rle_byte, U.y, nm1, lane + 32, le, e1);
    nm1 += __popc(U.y);
    uint32_t v2 = place_whole_chunk<kRle>(rec, pd + 64, DL, rle_byte, U.z, nm1, lane + 64, le, e2);
    nm1 += __popc(U.z);
    uint32_t v3 = place_whole_chunk<kRle>(rec, pd + 96, DL, rle_byte, U.w, nm1, lane + 96, le, e3);
    nm1 += __popc(U.w);
    const bool i0 = e0 <= lane, i1 = e1 <= lane + 32, i2 = e2 <= lane + 64, i3 = e3 <= lane + 96;
    const uint32_t o0 = __ballot_sync(kFull, i0), o1 = __ballot_sync(kFull, i1), o2 = __ballot_sync(kFull, i2), o3 = __ballot_sync(kFull, i3);
    if (o0 | o1 | o2 | o3) {
        uint32_t vp = 0;
        SZB_RESOLVE(0u, v0, i0, lane - e0, o0)
        SZB_RESOLVE(1u, v1, i1, lane + 32 - e1, o1)
        SZB_RESOLVE(2u, v2, i2, lane + 64 - e2, o2)
        SZB_RESOLVE(3u, v3, i3, lane + 96 - e3, o3)
    }
    pd[0] = (uint8_t)v0;
    pd[32] = (uint8_t)v1;
    pd[64] = (uint8_t)v2;
    pd[96] = (uint8_t)v3;
    __syncwarp();  // the line is in memory for the loads of the next one
}

// One chunk of a PARTIAL line (a block's first and last line; nothing in flight): all five steps at once.
//   u       the chunk's bitmap word (outside the part of the line being produced: cleared)
//   nm1     segments of the block that start below the chunk, minus 1
// Returns the byte (when its source is in memory); inl: its source is a byte of this step, at line position sr.
template <bool kRle>
__device__ __forceinline__ uint32_t place_chunk(const uint32_t *__restrict__ rec, const uint8_t *pd, long long DL, uint32_t rle_byte,
                                                uint32_t u, uint32_t nm1, uint32_t rel, uint32_t lo, bool live, uint32_t le, bool &inl,
                                                uint32_t &sr) {
    uint32_t idx = nm1 + __popc(u & le);  // the last segment that starts at or before my byte
    if (!live) idx = 0;
    const uint32_t e = rec[idx];
    // a match source at or above the line's first new byte is a byte of this step (a literal entry is >= 2^31: never)
    inl = live && e <= rel - lo;
    sr = rel - e;
    const bool is_l = e >= kRecLiteral;
    const uint8_t *p = pd - e;
    if (is_l) p += DL;
    uint32_t v = 0;
    if (kRle && is_l) {
        v = rle_byte;
    } else if (live && !inl) {
        v = *p;
    }
    return v;
}

// The bytes [lo, hi) of the line whose chunk-0 bytes are at pd.
template <bool kRle>
__device__ __forceinline__ void place_partial_line(const uint32_t *__restrict__ rec, uint8_t *pd, long long DL, uint32_t rle_byte, uint32_t lo,
                                                   uint32_t hi, uint4 U, uint32_t &nm1, uint32_t lane, uint32_t le) {
    bool i0, i1, i2, i3;
    uint32_t s0, s1, s2, s3;
    const bool a0 = lane >= lo && lane < hi, a1 = lane + 32 >= lo && lane + 32 < hi, a2 = lane + 64 >= lo && lane + 64 < hi,
               a3 = lane + 96 >= lo && lane + 96 < hi;
    uint32_t v0 = place_chunk<kRle>(rec, pd, DL, rle_byte, U.x, nm1, lane, lo, a0, le, i0, s0);
    nm1 += __popc(U.x);
    uint32_t v1 = place_chunk<kRle>(rec, pd + 32, DL, rle_byte, U.y, nm1, lane + 32, lo, a1, le, i1, s1);
    nm1 += __popc(U.y);
    uint32_t v2 = place_chunk<kRle>(rec, pd + 64, DL, rle_byte, U.z, nm1, lane + 64, lo, a2, le, i2, s2);
    nm1 += __popc(U.z);
    uint32_t v3 = place_chunk<kRle>(rec, pd + 96, DL, rle_byte, U.w, nm1, lane + 96, lo, a3, le, i3, s3);
    nm1 += __popc(U.w);
    const uint32_t o0 = __ballot_sync(kFull, i0), o1 = __ballot_sync(kFull, i1), o2 = __ballot_sync(kFull, i2), o3 = __ballot_sync(kFull, i3);
    if (o0 | o1 | o2 | o3) {
        uint32_t vp = 0;
        SZB_RESOLVE(0u, v0, i0, s0, o0)
        SZB_RESOLVE(1u, v1, i1, s1, o1)
        SZB_RESOLVE(2u, v2, i2, s2, o2)
        SZB_RESOLVE(3u, v3, i3, s3, o3)
    }
    if (a0) pd[0] = (uint8_t)v0;
    if (a1) pd[32] = (uint8_t)v1;
    if (a2) pd[64] = (uint8_t)v2;
    if (a3) pd[96] = (uint8_t)v3;
    __syncwarp();  // the line is in memory for the loads of the next one
}

#ifndef SZB_PLACE_WARPS
#define SZB_PLACE_WARPS 4
#endif
#ifndef SZB_PLACE_MIN_CTAS
#define SZB_PLACE_MIN_CTAS 10
#endif
constexpr int kPlaceWarps = SZB_PLACE_WARPS;

__device__ __forceinline__ uint32_t popc4(uint4 U) { return __popc(U.x) + __popc(U.y) + __popc(U.z) + __popc(U.w); }

// One block: q positions [bs, be) of the frame.  dq[x]: the output byte at q position x; litq[x - e]: the literal byte at q
// position x when e match bytes of the block lie in front of it.
template <bool kRle>
__device__ __forceinline__ void place_block(const uint32_t *__restrict__ rec, uint8_t *dq, const uint8_t *litq, uint32_t rle_byte,
                                            const uint32_t *bm, uint32_t bs, uint32_t be, uint32_t lane, uint32_t le) {
    uint32_t nm1 = 0xFFFFFFFFu;  // 0 - 1
    long long DL = (litq - dq) + (long long)kRecLiteral;  // a literal entry carries bit 31
    SZB_KEEP_PTR(rec);
    SZB_KEEP(DL);
    uint32_t X = bs & ~127u;
    if (bs != X || be < X + 128) {  // the block's first line, when it is not a whole one
        const uint32_t lo = bs - X, hi = be - X < 128 ? be - X : 128;
        uint4 U = *reinterpret_cast<const uint4 *>(bm + (X >> 5));
        U.x &= range_mask(lo, hi);
        U.y &= range_mask(lo > 32 ? lo - 32 : 0, hi > 32 ? hi - 32 : 0);
        U.z &= range_mask(lo > 64 ? lo - 64 : 0, hi > 64 ? hi - 64 : 0);
        U.w &= range_mask(lo > 96 ? lo - 96 : 0, hi > 96 ? hi - 96 : 0);
        uint8_t *pd = dq + X + lane;
        SZB_KEEP_PTR(pd);
        place_partial_line<kRle>(rec, pd, DL, rle_byte, lo, hi, U, nm1, lane, le);
        X += 128;
    }
    const uint32_t Xe = be & ~127u;  // whole lines: [X, Xe)
    if (X < Xe) {
        const uint4 *bq = reinterpret_cast<const uint4 *>(bm + (X >> 5));
        uint8_t *pd = dq + X + lane;
        SZB_KEEP_PTR(pd);
        uint4 U = *bq;
        for (; X < Xe; X += 128) {
            bq++;
            const uint4 Un = *bq;  // the next line's words are on their way (the bitmap covers one line more than the output)
            place_whole_line<kRle>(rec, pd, DL, rle_byte, U, nm1, lane, le);
            U = Un;
            pd += 128;
        }
    }
    if (X < be) {  // the block's last line, when it is not a whole one (lo = 0: X >= bs here)
        const uint32_t hi = be - X;
        uint4 U = *reinterpret_cast<const uint4 *>(bm + (X >> 5));
        U.x &= range_mask(0, hi);
        U.y &= range_mask(0, hi > 32 ? hi - 32 : 0);
        U.z &= range_mask(0, hi > 64 ? hi - 64 : 0);
        U.w &= range_mask(0, hi > 96 ? hi - 96 : 0);
        uint8_t *pd = dq + X + lane;
        SZB_KEEP_PTR(pd);
        place_partial_line<kRle>(rec, pd, DL, rle_byte, 0, hi, U, nm1, lane, le);
    }
}

// One warp per frame: frames exec_list[first_slot, first_slot + n_slots) that k_resolve took and found executable.
__global__ void __launch_bounds__(kPlaceWarps * 32, SZB_PLACE_MIN_CTAS) k_place(DeviceBatch a, uint32_t first_slot, uint32_t n_slots) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t slot = blockIdx.x * kPlaceWarps + (threadIdx.x >> 5);
    if (slot >= n_slots) return;
    const uint32_t f = a.exec_list[first_slot + slot];
    if (a.place_state[f] != SZB_OK || a.frame_status[f] != SZB_OK) return;  // k_resolve, k_frame_verdict
    const uint32_t b0 = a.frames[f].first_block, nb = a.frames[f].nblocks;
    if (nb == 0) return;
    const uint32_t A = (uint32_t)(reinterpret_cast<uintptr_t>(a.dst) & 127);
    const uint64_t frame_start = a.out_off[b0];
    const uint64_t line0 = (frame_start + A) & ~(uint64_t)127;
    const uint32_t *bm = a.bm + ((line0 >> 5) + 4ull * f);
    uint8_t *const dq = a.dst + line0 - A;  // 128-byte aligned
    const uint32_t le = 0xFFFFFFFFu >> (31 - lane);
    for (uint32_t bi = 0; bi < nb; bi++) {
        const uint32_t b = b0 + bi;
        const szb_block_desc *d = a.blocks + b;
        if (d->type != 2 || d->nseq == 0) continue;  // written by k_execute_bodies already
        const uint32_t bs = (uint32_t)(a.out_off[b] + A - line0), be = bs + (uint32_t)a.out_size[b];
        const uint32_t lit_type = d->lit_type;
        const uint8_t *payload = a.src + d->src_off;
        const uint32_t *rec = a.rec + a.rec_off[b];
        if (lit_type == 1) {  // RLE literals: every literal byte is that byte
            place_block<true>(rec, dq, dq, payload[d->lit_hdr_bytes], bm, bs, be, lane, le);
        } else {
            const uint8_t *lit = lit_type == 0 ? payload + d->lit_hdr_bytes : a.litbuf + d->lit_buf_off;
            place_block<false>(rec, dq, lit - bs, 0, bm, bs, be, lane, le);
        }
    }
}
