// warpsim.h -- TEST INFRASTRUCTURE: a minimal CPU emulation of one CUDA warp, enough to run the warp-cooperative kernels
// of execute_long.cuh (full-mask shuffles, ballots, votes; no shared memory, no __syncthreads) under pytest -m "not gpu".
//
// Every lane of a warp is a coroutine (ucontext).  A lane runs until it reaches a warp collective, leaves its operands
// in its slot and yields; when all live lanes have arrived at the same collective the scheduler computes every lane's
// result and resumes them.  Warps and CTAs run one after the other, in the order the caller chooses (kernels whose result
// must not depend on scheduling are run in several orders).  Never linked into libszb200.so.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <ucontext.h>

#include <cstdio>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

namespace warpsim {

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
};
inline Dim3 threadIdx, blockIdx, blockDim, gridDim;

enum Op { OP_NONE, OP_SHFL, OP_SHFL_UP, OP_BALLOT };
struct Lane {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    Op op = OP_NONE;
    uint64_t val = 0;
    uint32_t arg = 0;
    uint64_t res = 0;
};
inline Lane lanes[32];
inline ucontext_t sched_ctx;
inline int cur = 0;
inline void (*warp_body)() = nullptr;
inline uint64_t collectives = 0;

inline void trampoline() {
    warp_body();
    lanes[cur].done = true;
    swapcontext(&lanes[cur].ctx, &sched_ctx);
}

inline uint64_t collective(Op op, uint64_t val, uint32_t arg) {
    Lane &l = lanes[cur];
    l.op = op;
    l.val = val;
    l.arg = arg;
    swapcontext(&l.ctx, &sched_ctx);
    return lanes[cur].res;
}

// runs the 32 lanes of warp `warp` of the current CTA (threadIdx.x = warp * 32 + lane)
inline void run_warp(unsigned warp) {
    for (int i = 0; i < 32; i++) {
        Lane &l = lanes[i];
        if (l.stack.empty()) l.stack.resize(256 * 1024);
        l.done = false;
        l.op = OP_NONE;
        getcontext(&l.ctx);
        l.ctx.uc_stack.ss_sp = l.stack.data();
        l.ctx.uc_stack.ss_size = l.stack.size();
        l.ctx.uc_link = nullptr;
        makecontext(&l.ctx, trampoline, 0);
    }
    for (;;) {
        int live = 0;
        for (int i = 0; i < 32; i++) {
            if (lanes[i].done) continue;
            cur = i;
            threadIdx.x = warp * 32 + i;
            lanes[i].op = OP_NONE;
            swapcontext(&sched_ctx, &lanes[i].ctx);
            if (!lanes[i].done) live++;
        }
        if (!live) return;
        Op op = OP_NONE;
        for (int i = 0; i < 32; i++) {
            if (lanes[i].done) continue;
            if (op == OP_NONE) op = lanes[i].op;
            if (lanes[i].op != op || op == OP_NONE) {
                fprintf(stderr, "warpsim: lanes diverged at a full-mask collective\n");
                abort();
            }
        }
        collectives++;
        if (op == OP_BALLOT) {
            uint32_t m = 0;
            for (int i = 0; i < 32; i++)
                if (!lanes[i].done && lanes[i].val) m |= 1u << i;
            for (int i = 0; i < 32; i++) lanes[i].res = m;
        } else if (op == OP_SHFL) {
            for (int i = 0; i < 32; i++) lanes[i].res = lanes[lanes[i].arg & 31].val;
        } else {  // OP_SHFL_UP
            for (int i = 0; i < 32; i++) lanes[i].res = (uint32_t)i >= lanes[i].arg ? lanes[i - lanes[i].arg].val : lanes[i].val;
        }
    }
}

// launches `body` (a wrapper that calls the kernel with its arguments) over the grid; cta_order may permute the CTAs
template <class F>
inline void launch(unsigned grid, unsigned block, F body, const std::vector<unsigned> *cta_order = nullptr) {
    static F *fn;
    fn = &body;
    warp_body = [] { (*fn)(); };
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned c = 0; c < grid; c++) {
        blockIdx.x = cta_order ? (*cta_order)[c] : c;
        for (unsigned w = 0; w < (block + 31) / 32; w++) run_warp(w);
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
inline uint32_t __ballot_sync(uint32_t, bool p) { return (uint32_t)warpsim::collective(warpsim::OP_BALLOT, p, 0); }
inline bool __any_sync(uint32_t, bool p) { return warpsim::collective(warpsim::OP_BALLOT, p, 0) != 0; }
inline void __syncwarp() { warpsim::collective(warpsim::OP_BALLOT, 0, 0); }
inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline uint32_t __ldcg(const uint32_t *p) { return *(const volatile uint32_t *)p; }
inline void __stcg(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }
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
