// warpsim.h -- TEST INFRASTRUCTURE: a minimal CPU emulation of one CUDA thread block, enough to run the warp-cooperative
// kernels of stage 4 (execute.cuh, execute_long.cuh: full-mask shuffles, ballots, votes, warp reductions, __syncwarp,
// __syncthreads, static shared memory, shared-memory atomics) under pytest -m "not gpu".
//
// Every thread of a CTA is a coroutine (ucontext).  A thread runs until it reaches a collective, leaves its operands in its
// slot and yields.  A warp-level collective is resolved when all live lanes of that warp have arrived at it; __syncthreads
// when all live threads of the CTA have.  CTAs run one after the other, in the order the caller chooses (kernels whose
// result must not depend on scheduling are run in several orders); static __shared__ variables are therefore plain statics.
// Global and shared memory are the host's memory: sequentially consistent, which is stronger than the device -- the
// emulator checks algorithms and index arithmetic, not memory-ordering bugs (compute-sanitizer runs cover those on the GPU).
// Never linked into libszb200.so.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <ucontext.h>

#include <cstdio>
#include <vector>

#define SZB_WARPSIM 1
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct alignas(16) uint4 {
    uint32_t x, y, z, w;
};
struct alignas(8) uint2 {
    uint32_t x, y;
};
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace warpsim {

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
};
inline Dim3 threadIdx, blockIdx, blockDim, gridDim;

enum Op { OP_NONE, OP_SHFL, OP_SHFL_UP, OP_BALLOT, OP_REDUCE_MAX, OP_CTA_BARRIER, OP_NAMED_BARRIER };
constexpr int kMaxThreads = 1024;
struct Thread {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    Op op = OP_NONE;
    uint64_t val = 0;
    uint32_t arg = 0;
    uint64_t res = 0;
};
inline Thread threads[kMaxThreads];
inline ucontext_t sched_ctx;
inline int cur = 0;
inline void (*cta_body)() = nullptr;
inline uint64_t collectives = 0;

inline void trampoline() {
    cta_body();
    threads[cur].done = true;
    swapcontext(&threads[cur].ctx, &sched_ctx);
}

inline uint64_t collective(Op op, uint64_t val, uint32_t arg) {
    Thread &t = threads[cur];
    t.op = op;
    t.val = val;
    t.arg = arg;
    swapcontext(&t.ctx, &sched_ctx);
    return threads[cur].res;
}

[[noreturn]] inline void die(const char *what) {
    fprintf(stderr, "warpsim: %s\n", what);
    abort();
}

// resolves the collective the live lanes of warp w wait at; false when the warp waits at the CTA barrier (or is done)
inline bool resolve_warp(int w) {
    Thread *l = threads + 32 * w;
    Op op = OP_NONE;
    for (int i = 0; i < 32; i++) {
        if (l[i].done) continue;
        if (op == OP_NONE) op = l[i].op;
        if (l[i].op != op || op == OP_NONE) die("lanes of a warp diverged at a full-mask collective");
    }
    if (op == OP_NONE || op == OP_CTA_BARRIER || op == OP_NAMED_BARRIER) return false;
    collectives++;
    if (op == OP_BALLOT) {
        uint32_t m = 0;
        for (int i = 0; i < 32; i++)
            if (!l[i].done && l[i].val) m |= 1u << i;
        for (int i = 0; i < 32; i++) l[i].res = m;
    } else if (op == OP_REDUCE_MAX) {
        uint64_t m = 0;
        for (int i = 0; i < 32; i++)
            if (!l[i].done && l[i].val > m) m = l[i].val;
        for (int i = 0; i < 32; i++) l[i].res = m;
    } else if (op == OP_SHFL) {
        for (int i = 0; i < 32; i++) l[i].res = l[l[i].arg & 31].val;
    } else {  // OP_SHFL_UP
        for (int i = 0; i < 32; i++) l[i].res = (uint32_t)i >= l[i].arg ? l[i - l[i].arg].val : l[i].val;
    }
    return true;
}

// runs one CTA of `nthreads` threads to completion
inline void run_cta(unsigned nthreads) {
    if (nthreads > kMaxThreads || nthreads % 32) die("CTA size must be a multiple of 32, at most 1024");
    const int nwarps = (int)nthreads / 32;
    for (unsigned i = 0; i < nthreads; i++) {
        Thread &t = threads[i];
        if (t.stack.empty()) t.stack.resize(96 * 1024);
        t.done = false;
        t.op = OP_NONE;
        getcontext(&t.ctx);
        t.ctx.uc_stack.ss_sp = t.stack.data();
        t.ctx.uc_stack.ss_size = t.stack.size();
        t.ctx.uc_link = nullptr;
        makecontext(&t.ctx, trampoline, 0);
    }
    std::vector<char> runnable(nwarps, 1);  // the warp's lanes hold results (or have not started) and can run on
    for (;;) {
        bool progressed = false;
        for (int w = 0; w < nwarps; w++) {
            while (runnable[w]) {  // a warp runs from collective to collective until it meets the CTA barrier or ends
                for (int i = 0; i < 32; i++) {
                    Thread &t = threads[32 * w + i];
                    if (t.done) continue;
                    cur = 32 * w + i;
                    threadIdx.x = (unsigned)cur;
                    t.op = OP_NONE;
                    swapcontext(&sched_ctx, &t.ctx);
                }
                progressed = true;
                runnable[w] = resolve_warp(w);
            }
        }
        // every warp is done or waits at a barrier.  A named barrier (bar.sync id, nthreads: arg = id << 16 | nthreads) opens
        // when that many threads wait at it; the warps at the CTA barrier stay where they are.
        {
            bool opened = false;
            for (int w = 0; w < nwarps && !opened; w++) {
                const Thread &t0 = threads[32 * w];
                if (t0.done || t0.op != OP_NAMED_BARRIER) continue;
                unsigned waiting = 0;
                for (unsigned i = 0; i < nthreads; i++)
                    if (!threads[i].done && threads[i].op == OP_NAMED_BARRIER && threads[i].arg == t0.arg) waiting++;
                if (waiting > (t0.arg & 0xFFFF)) die("more threads at a named barrier than it was declared for");
                if (waiting == (t0.arg & 0xFFFF)) {
                    for (int v = 0; v < nwarps; v++)
                        if (!threads[32 * v].done && threads[32 * v].op == OP_NAMED_BARRIER && threads[32 * v].arg == t0.arg) runnable[v] = 1;
                    opened = true;
                }
            }
            if (opened) {
                collectives++;
                continue;
            }
        }
        bool any_live = false;
        for (unsigned i = 0; i < nthreads; i++) {
            if (threads[i].done) continue;
            any_live = true;
            if (threads[i].op == OP_NAMED_BARRIER) die("deadlock at a named barrier");
            if (threads[i].op != OP_CTA_BARRIER) die("a thread waits at a warp collective its warp cannot complete");
        }
        if (!any_live) return;
        if (!progressed) die("deadlock at __syncthreads");
        collectives++;
        for (int w = 0; w < nwarps; w++) {
            bool live = false;
            for (int i = 0; i < 32; i++) live |= !threads[32 * w + i].done;
            runnable[w] = live;
        }
    }
}

// launches `body` (a wrapper that calls the kernel with its arguments) over the grid; cta_order may permute the CTAs
template <class F>
inline void launch(unsigned grid, unsigned block, F body, const std::vector<unsigned> *cta_order = nullptr) {
    static F *fn;
    fn = &body;
    cta_body = [] { (*fn)(); };
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned c = 0; c < grid; c++) {
        blockIdx.x = cta_order ? (*cta_order)[c] : c;
        run_cta((block + 31) / 32 * 32);
    }
}

}  // namespace warpsim

using warpsim::blockDim;
using warpsim::blockIdx;
using warpsim::gridDim;
using warpsim::threadIdx;

template <class T, class S>
inline T __shfl_sync(uint32_t, T v, S src) { return (T)warpsim::collective(warpsim::OP_SHFL, (uint64_t)v, (uint32_t)src); }
template <class T, class S>
inline T __shfl_up_sync(uint32_t, T v, S d) { return (T)warpsim::collective(warpsim::OP_SHFL_UP, (uint64_t)v, (uint32_t)d); }
template <class T>
inline T __shfl_xor_sync(uint32_t, T v, int mask) { return (T)warpsim::collective(warpsim::OP_SHFL, (uint64_t)v, (uint32_t)((warpsim::cur & 31) ^ mask)); }
inline uint32_t __ballot_sync(uint32_t, bool p) { return (uint32_t)warpsim::collective(warpsim::OP_BALLOT, p, 0); }
inline bool __any_sync(uint32_t, bool p) { return warpsim::collective(warpsim::OP_BALLOT, p, 0) != 0; }
inline uint32_t __reduce_max_sync(uint32_t, uint32_t v) { return (uint32_t)warpsim::collective(warpsim::OP_REDUCE_MAX, v, 0); }
inline void __syncwarp() { warpsim::collective(warpsim::OP_BALLOT, 0, 0); }
inline void __syncthreads() { warpsim::collective(warpsim::OP_CTA_BARRIER, 0, 0); }
// bar.sync id, nthreads: whole warps only (every lane of a warp arrives together)
inline void __named_barrier(uint32_t id, uint32_t nthreads) { warpsim::collective(warpsim::OP_NAMED_BARRIER, 0, (id << 16) | nthreads); }
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
    const uint64_t t = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((t >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31)); }
inline uint32_t __ldcg(const uint32_t *p) { return *(const volatile uint32_t *)p; }
inline void __stcg(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }
inline uint32_t atomicOr(uint32_t *p, uint32_t v) {
    const uint32_t old = *p;
    *p = old | v;
    return old;
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    const unsigned long long old = *p;
    *p = old + v;
    return old;
}
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
    const unsigned long long old = *p;
    if (v < old) *p = v;
    return old;
}
