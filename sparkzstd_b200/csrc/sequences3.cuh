// sequences3.cuh -- stage 3 with THREE LANES PER BLOCK (sm_100a): k_decode_sequences3.
// Included by kernels.cuh inside namespace szb, after k_decode_sequences (whose ring helpers and slow path it shares).
//
// replaces: structure/sequences.go:126-206 DecodeSequences, :64-123 DecodeSequence (same as k_decode_sequences)
//
// k_decode_sequences gives a block to ONE lane, which walks its three FSE states itself: ~96 warp instructions and ~280 cycles
// per step of a warp's 21 chains, one warp per scheduler (shared memory -- 2.7 KB of tables and bit ring per chain -- decides
// how many chains an SM holds, not how many warps).  Lanes per warp do not matter (profiles/r02r_*); what bounds the kernel is
// chains in flight x the latency of a step.  Here a block gets three lanes, one per state (LL, ML, OF): each loads ONE cell,
// finds ITS field widths, the three exchange them with shuffles inside the triple, and each lane cuts its two fields (its
// extra bits, its state bits) out of the same 64-bit window with two independent shifts instead of peeling six fields off
// one after the other.  A step is ~55 instructions and a chain of ~130 cycles; a warp carries 10 chains, an SM eight such
// warps (two per scheduler), i.e. 80 chains: the same chains in flight, each advancing about twice as fast.
//
// The warp stays whole: every lane runs as many groups as the warp's longest block has (idle rounds change nothing), so that the
// shuffles are plain full-mask shuffles -- a shuffle over a sub-warp mask costs five more instructions -- and everything that
// depends on a lane's role is arithmetic with per-lane constants, not a branch (a branch on the role splits the warp in three).
//
// What stays as it is: the 16-bit resident cells, the 128-byte bit ring per chain topped up with cp.async once per group of four
// sequences (issued by the LL lane, visible to the triple after a __syncwarp over the triple), the speculative group with its
// `bad` flag and the byte-wise redo (every lane of the triple redoes the group with all three states: rare), the three SoA
// output arrays (each lane stores its own field, 16 bytes per group).
#pragma once

constexpr int kSeq3Chains = 10;                                              // blocks per warp (30 lanes busy)
constexpr uint32_t kSeq3TabBytes = kSeq3Chains * kTabSlotWords * 2;         // u16 cells
constexpr uint32_t kSeq3RingOff = kSeq3TabBytes;                             // byte offset of the rings
constexpr uint32_t kSeq3LutWord = (kSeq3RingOff + kSeq3Chains * kSeqRingStride) / 4;  // ll[64] | ml[64] | (unused)[64]
constexpr uint32_t kSeq3SmemBytes = (kSeq3LutWord + 192) * 4;                // 27.8 KB: eight one-warp CTAs per SM

// n bits (n <= 32, possibly 0) of the 64-bit window hi:lo, starting `off` bits below its top (off + n <= 64)
__device__ __forceinline__ uint32_t seq3_take(uint32_t hi, uint32_t lo, uint32_t off, uint32_t n) {
    const uint32_t t = off < 32 ? __funnelshift_l(lo, hi, off) : (lo << (off & 31));  // the top 32 bits of (window << off)
    return (t >> 1) >> (31 - n);
}

struct Seq3Lane {
    uint32_t s;   // my FSE state
    int32_t pos;  // stream bits not consumed yet (the same in the three lanes of a block)
};

// what a lane's role means, as masks and constants (role 0: literal lengths, 1: match lengths, 2: offsets)
struct Seq3Role {
    uint32_t xo_of, xo_ml;  // my extra bits lie behind the offsets' (roles 0, 1) and the match lengths' (role 0) extra bits
    uint32_t so_ll, so_ml;  // my state bits lie behind the literal lengths' (roles 1, 2) and the match lengths' (role 2) state bits
    uint32_t of_one;        // 1 for the offsets lane: its base value is 1 << code (sequences.go:99-104)
    uint32_t lut;           // word index of my 64-entry table: base | extra bits << 24 (offsets: 0 | code << 24)
};

// One sequence: my field of it and my state's update (upd: all ones, or 0 for a block's last sequence, sequences.go:178).
__device__ __forceinline__ void seq3_step(const uint32_t *sw, const uint16_t *mytab, const uint8_t *ring, int32_t sp_mis, uint32_t my_al,
                                          const Seq3Role &R, uint32_t lane0, uint32_t upd, Seq3Lane &L, int32_t &bad, uint32_t &v) {
    const uint32_t c = mytab[L.s];  // peek my cell (sequences.go:67-78)
    const uint32_t code = c & 63;
    const uint32_t u = sw[R.lut + code];
    const uint32_t extra = u >> 24;
    const uint32_t base = (u & 0xFFFFFF) + (R.of_one << code);
    const uint32_t nb = (my_al - bfind(c >> 6)) & upd;  // NumberOfBits = AL - highbit(next) (fse.go:212)
    const uint32_t pack = extra | (nb << 8);
    const uint32_t p_ll = __shfl_sync(kFull, pack, lane0), p_ml = __shfl_sync(kFull, pack, lane0 + 1), p_of = __shfl_sync(kFull, pack, lane0 + 2);
    const uint32_t e_ll = p_ll & 255, e_ml = p_ml & 255, e_of = p_of & 255;
    const uint32_t n_ll = p_ll >> 8, n_ml = p_ml >> 8, n_of = p_of >> 8;
    const uint32_t E = e_of + e_ml + e_ll;
    const uint32_t total = E + n_ll + n_ml + n_of;
    // the stream holds, from the top: OF extra, ML extra, LL extra, then LL state, ML state, OF state (sequences.go:99-120,178-194)
    const uint32_t xo = (e_of & R.xo_of) + (e_ml & R.xo_ml);
    const uint32_t so = E + (n_ll & R.so_ll) + (n_ml & R.so_ml);
    // window: the 8 bytes ending at the byte that holds bit pos-1, that bit moved to bit 63 (as fast_step)
    const int32_t top = sp_mis + ((L.pos - 1) >> 3);
    const uint32_t ap = (uint32_t)(top - 7);
    const uint32_t mis = ap & 3;
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(ring + ((ap - mis) & (kSeqRingBytes - 1)));
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    uint32_t lo = __funnelshift_r(w0, w1, mis * 8), hi = __funnelshift_r(w1, w2, mis * 8);
    const uint32_t k = 7 - ((uint32_t)(L.pos - 1) & 7);
    hi = __funnelshift_l(lo, hi, k);
    lo <<= k;
    const uint32_t x = seq3_take(hi, lo, xo, extra);
    const uint32_t b = seq3_take(hi, lo, so, nb);
    L.pos -= (int32_t)total;
    bad |= (57 - (int32_t)total) | L.pos;
    v = base + x;
    const uint32_t next = ((c >> 6) << nb) - (1u << my_al) + b;  // Baseline = (next << nb) - 2^AL (fse.go:213)
    L.s = upd ? next : L.s;
}

// ring upkeep: the copies are issued by the triple's first lane only; every lane keeps the same books
__device__ __forceinline__ void ring3_fill(bool issue, uint32_t ring_saddr, const uint4 *chunk0, int32_t top, int32_t &lowreq) {
    int32_t need = (top - kSeqRingAhead) >> 4;
    need = need < 0 ? 0 : need;
    while (lowreq > need) ring_fetch_if(issue, ring_saddr, chunk0, --lowreq);
}
__device__ __forceinline__ void ring3_topup(bool act, bool issue, uint32_t ring_saddr, const uint4 *chunk0, int32_t top, int32_t &lowreq) {
    int32_t need = (top - kSeqRingAhead) >> 4;
    need = need < 0 ? 0 : need;
    const bool p1 = act && lowreq - 1 >= need, p2 = act && lowreq - 2 >= need;
    ring_fetch_if(p1 && issue, ring_saddr, chunk0, lowreq - 1);
    ring_fetch_if(p2 && issue, ring_saddr, chunk0, lowreq - 2);
    lowreq -= (int32_t)p1 + (int32_t)p2;
}

__global__ void __launch_bounds__(32) k_decode_sequences3(DeviceBatch a) {
    extern __shared__ __align__(16) uint32_t sw[];
    const uint32_t lane = threadIdx.x;
    const uint32_t chain = lane / 3, role = lane - 3 * chain;  // role 0: LL, 1: ML, 2: OF
    const uint32_t first = blockIdx.x * kSeq3Chains;
    const uint32_t n_here = a.n_seq - first < (uint32_t)kSeq3Chains ? a.n_seq - first : (uint32_t)kSeq3Chains;
    uint16_t *tabs = reinterpret_cast<uint16_t *>(sw);

    {   // tables: HBM arena -> shared memory, 2560 contiguous bytes per block, all copies in flight at once
        const uint32_t tabs_saddr = (uint32_t)__cvta_generic_to_shared(tabs);
        const uint8_t *arena = reinterpret_cast<const uint8_t *>(a.seq_tabs + (size_t)first * kTabSlotWords);
        const uint32_t chunks = n_here * (kTabSlotWords * 2 / 16);
        for (uint32_t c = lane; c < chunks; c += 32) cp_async16(tabs_saddr + c * 16, arena + (size_t)c * 16);
        asm volatile("cp.async.commit_group;");
    }
    for (uint32_t i = lane; i < 64; i += 32) {
        sw[kSeq3LutWord + i] = kLLBaseDev[i] | ((uint32_t)kLLExtraDev[i] << 24);
        sw[kSeq3LutWord + 64 + i] = kMLBaseDev[i] | ((uint32_t)kMLExtraDev[i] << 24);
        sw[kSeq3LutWord + 128 + i] = i << 24;  // offsets: as many extra bits as the code says, the base (1 << code) is added apart
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    // A lane without a block (lanes 30 and 31, a short last CTA, a block whose tables failed) stays with the warp and idles on
    // block 0's tables: nothing it computes is kept.
    bool alive = chain < n_here;
    const uint32_t ch = alive ? chain : 0u;
    const uint32_t lane0 = alive ? 3 * chain : lane;  // (an idle lane shuffles with itself)
    const uint32_t w = first + ch;
    const uint32_t b = a.seq_list[w];
    if (alive && a.seq_status[b] != SZB_OK) alive = false;  // its tables failed to build
    const szb_block_desc d = a.blocks[b];
    if (alive && (d.flags & SZB_BLOCK_TABLES_ONLY)) {  // a dictionary's row: its tables were built, it has no sequences to decode
        if (role == 0) a.out_size[b] = 0;
        alive = false;
    }
    const SeqInfo info = a.seq_info[w];
    const uint16_t *tll = tabs + ch * kTabSlotWords;
    const uint16_t *mytab = role == 0 ? tll : (role == 1 ? tll + 512 : tll + 1024);
    const uint32_t al_ll = info.al_ll, al_ml = info.al_ml, al_of = info.al_of;
    const uint32_t my_al = role == 0 ? al_ll : (role == 1 ? al_ml : al_of);
    Seq3Role R;
    R.xo_of = role != 2 ? ~0u : 0u;
    R.xo_ml = role == 0 ? ~0u : 0u;
    R.so_ll = role != 0 ? ~0u : 0u;
    R.so_ml = role == 2 ? ~0u : 0u;
    R.of_one = role == 2 ? 1u : 0u;
    R.lut = kSeq3LutWord + (role << 6);
    const uint32_t hdr = d.seq_off + d.seq_hdr_bytes + info.stream_off;
    const uint8_t *sp = a.src + d.src_off + hdr;
    const uint32_t len = alive ? d.block_size - hdr : 0u;

    // padding: zero bits then the first 1 bit, at most 8 (sequences.go:131-143)
    if (alive && (len == 0 || sp[len - 1] == 0)) {
        if (role == 0) a.seq_status[b] = SZB_ERR_BAD_PADDING;
        alive = false;
    }
    Seq3Lane L;
    L.s = 0;
    L.pos = 64;
    if (alive) {  // InitState in the order LL, OF, ML (sequences.go:145-159): every lane reads its own
        L.pos = (int32_t)(len * 8) - (__clz((uint32_t)sp[len - 1]) - 24 + 1);
        const int32_t my_at = L.pos - (int32_t)(role == 0 ? 0u : (role == 2 ? al_ll : al_ll + al_of));
        L.s = slow_read_bits(sp, my_at, my_al);
        L.pos -= (int32_t)(al_ll + al_of + al_ml);
    }

    const uint8_t *ring = reinterpret_cast<const uint8_t *>(sw) + kSeq3RingOff + ch * kSeqRingStride;
    const uint32_t ring_saddr = (uint32_t)__cvta_generic_to_shared(ring);
    const int32_t sp_mis = (int32_t)(reinterpret_cast<uintptr_t>(sp) & 15);
    const uint4 *chunk0 = reinterpret_cast<const uint4 *>(sp - sp_mis);
    const bool issue = alive && role == 0;
    int32_t lowreq = ((sp_mis + (int32_t)len - 1) >> 4) + 1;  // lowest chunk requested so far
    if (alive) ring3_fill(issue, ring_saddr, chunk0, sp_mis + ((L.pos - 1) >> 3), lowreq);
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    const uint32_t tmask = 7u << (3 * ch);  // only for the rare byte-wise redo, which a triple takes on its own
    const uint32_t *const lut_at = sw + ((int32_t)kSeq3LutWord - (int32_t)kSeqLutWord);  // slow_step looks its tables up at kSeqLutWord
    const uint32_t nseq = alive ? d.nseq : 0u;
    uint32_t *const my_out = (role == 0 ? a.seq_ll : (role == 1 ? a.seq_ml : a.seq_of)) + d.seq_buf_off;
    uint64_t sum = 0;  // of my field: the match lengths' sum makes the block's size
    const uint32_t n_upd = nseq ? nseq - 1 : 0u;  // every sequence but the last updates the states (sequences.go:178)
    const uint32_t my_groups = n_upd >> 2;
    const uint32_t max_groups = __reduce_max_sync(kFull, my_groups);
    int32_t low1 = lowreq, low2 = lowreq;  // lowreq one and two top-ups ago
    for (uint32_t g = 0; g < max_groups; g++) {
        const bool act = g < my_groups;
        const int32_t landed = low2 ? (low2 << 4) : -64;  // what was requested two top-ups ago has landed once wait_group 2 returns
        const int32_t top = sp_mis + ((L.pos - 1) >> 3);
        ring3_topup(act, issue, ring_saddr, chunk0, top, lowreq);
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();  // the first lane's landed copies are the triple's
        low2 = act ? low1 : low2;
        low1 = act ? lowreq : low1;
        const Seq3Lane S = L;
        int32_t bad = top - kSeqGroupReach - landed;  // the same number in all three lanes
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) seq3_step(sw, mytab, ring, sp_mis, my_al, R, lane0, ~0u, L, bad, v[j]);
        if (!act) L = S;  // an idle round changes nothing
        if (act && bad < 0) {  // redo the group byte-wise, every lane with all three states (rare)
            SeqLane T;
            T.s_ll = __shfl_sync(tmask, S.s, lane0);
            T.s_ml = __shfl_sync(tmask, S.s, lane0 + 1);
            T.s_of = __shfl_sync(tmask, S.s, lane0 + 2);
            T.pos = S.pos;
            uint32_t o[12];
            for (int j = 0; j < 4; j++) slow_step<true>(lut_at, tll, tll + 512, tll + 1024, sp, al_ll, al_ml, al_of, T, o[j], o[4 + j], o[8 + j]);
            L.s = role == 0 ? T.s_ll : (role == 1 ? T.s_ml : T.s_of);
            L.pos = T.pos;
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = role == 0 ? o[j] : (role == 1 ? o[4 + j] : o[8 + j]);
        }
        if (act) {
            *reinterpret_cast<uint4 *>(my_out + 4 * g) = make_uint4(v[0], v[1], v[2], v[3]);  // seq_buf_off is a multiple of 32 entries
            sum += (uint64_t)v[0] + v[1] + v[2] + v[3];
        }
    }
    {   // the last (at most four) sequences one at a time; the ring is made to cover them all at once
        if (alive) ring3_fill(issue, ring_saddr, chunk0, sp_mis + ((L.pos - 1) >> 3), lowreq);
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const int32_t landed = lowreq ? (lowreq << 4) : -64;
        const uint32_t i0 = 4 * my_groups;
        const uint32_t my_tail = nseq - i0;  // 1 .. 4 (0 for an idle lane)
        const uint32_t max_tail = __reduce_max_sync(kFull, my_tail);
        for (uint32_t t = 0; t < max_tail; t++) {
            const uint32_t i = i0 + t;
            const bool act = t < my_tail;
            const bool upd = i < n_upd;
            uint32_t v;
            const Seq3Lane S = L;
            int32_t bad = sp_mis + ((L.pos - 1) >> 3) - 10 - landed;
            seq3_step(sw, mytab, ring, sp_mis, my_al, R, lane0, upd ? ~0u : 0u, L, bad, v);
            if (!act) L = S;
            if (act && bad < 0) {
                SeqLane T;
                T.s_ll = __shfl_sync(tmask, S.s, lane0);
                T.s_ml = __shfl_sync(tmask, S.s, lane0 + 1);
                T.s_of = __shfl_sync(tmask, S.s, lane0 + 2);
                T.pos = S.pos;
                uint32_t o[3];
                if (upd)
                    slow_step<true>(lut_at, tll, tll + 512, tll + 1024, sp, al_ll, al_ml, al_of, T, o[0], o[1], o[2]);
                else
                    slow_step<false>(lut_at, tll, tll + 512, tll + 1024, sp, al_ll, al_ml, al_of, T, o[0], o[1], o[2]);
                L.s = role == 0 ? T.s_ll : (role == 1 ? T.s_ml : T.s_of);
                L.pos = T.pos;
                v = role == 0 ? o[0] : (role == 1 ? o[1] : o[2]);
            }
            if (act) {
                my_out[i] = v;
                sum += v;
            }
        }
    }
    if (!alive) return;
    // the stream must be consumed exactly (sequences.go:197-204)
    if (role == 0) a.seq_status[b] = L.pos == 0 ? SZB_OK : SZB_ERR_NOT_ALL_BITS_USED;
    if (role == 1) a.out_size[b] = (uint64_t)d.lit_regen + sum;
}
