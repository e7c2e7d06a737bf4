// cuda_emu.h -- TEST/DEBUG HARNESS ONLY: lets g++ compile the kernel translation unit (csrc/*.cuh) for the host so
// the warp-synchronous code can be single-stepped and checked against the oracle in a container without a GPU.
// One warp = 32 cooperative fibers (ucontext) on one OS thread; a fiber yields only inside a warp collective
// (__syncwarp / __shfl_* / __ballot_sync / __any_sync), which is exactly where real lanes exchange data.
// This is not a product path: nothing under av_aloha_b200/ includes it and libavsim.so has no host fallback.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__
#define AV_SHARED static
#define __grid_constant__
#define __constant__ static const
static inline float __fdividef(float a, float b) { return a / b; }
#define __align__(n) alignas(n)

struct float4 {
    float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
struct dim3_emu {
    int x, y, z;
};

namespace emu {
struct Warp {
    ucontext_t main, ctx[32];
    char *stack[32];
    bool done[32];
    int cur;
    int arrived;
    uint64_t gen;
    uint32_t xchg[2][32];
    std::function<void()> body;
};
extern Warp W;
extern dim3_emu g_blockIdx, g_gridDim, g_blockDim;
static inline void yield_() { swapcontext(&W.ctx[W.cur], &W.main); }
static inline const uint32_t *exchange(uint32_t v) {
    int lane = W.cur;
    uint64_t my = W.gen;
    W.xchg[my & 1][lane] = v;
    if (++W.arrived == 32) { W.arrived = 0; W.gen++; }
    else while (W.gen == my) yield_();
    return W.xchg[my & 1];
}
void run_block(int block, int grid, const std::function<void()> &body);
}  // namespace emu

struct threadIdx_t { int y = 0, z = 0; struct { operator int() const { return emu::W.cur; } } x; };
static threadIdx_t threadIdx;
#define blockIdx emu::g_blockIdx
#define gridDim emu::g_gridDim
#define blockDim emu::g_blockDim

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline unsigned __float_as_uint(float f) { return f2u(f); }
static inline float __uint_as_float(unsigned u) { return u2f(u); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::exchange(0); }
static inline void __syncthreads() { emu::exchange(0); }   // one warp per emulated block
static inline int __syncthreads_or(int pred) {
    const uint32_t *x = emu::exchange(pred ? 1u : 0u);
    int r = 0;
    for (int i = 0; i < 32; i++) r |= x[i] != 0;
    return r;
}
static inline int __shfl_sync(unsigned, int v, int src) { return (int)emu::exchange((uint32_t)v)[src & 31]; }
static inline float __shfl_sync(unsigned, float v, int src) { return u2f(emu::exchange(f2u(v))[src & 31]); }
static inline float __shfl_sync(unsigned, float v, int src, int width) {
    int l = emu::W.cur, base = l & ~(width - 1);
    return u2f(emu::exchange(f2u(v))[base + (src & (width - 1))]);
}
static inline int __shfl_xor_sync(unsigned, int v, int o) { int l = emu::W.cur; return (int)emu::exchange((uint32_t)v)[(l ^ o) & 31]; }
static inline float __shfl_xor_sync(unsigned, float v, int o) { int l = emu::W.cur; return u2f(emu::exchange(f2u(v))[(l ^ o) & 31]); }
static inline int __shfl_up_sync(unsigned, int v, int off) {
    int l = emu::W.cur;
    const uint32_t *x = emu::exchange((uint32_t)v);
    return l >= off ? (int)x[l - off] : v;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    const uint32_t *x = emu::exchange(pred ? 1u : 0u);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (x[i] ? 1u : 0u) << i;
    return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    const uint32_t *x = emu::exchange(v);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r = x[i] > r ? x[i] : r;
    return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    const uint32_t *x = emu::exchange(v);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; i++) r = x[i] < r ? x[i] : r;
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline long long clock64() { return 0; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }   // fibers never preempt between collectives
using std::isfinite;
using std::max;
using std::min;
