// avsim_math.cuh -- small fp32 vector helpers + warp collectives used by the step kernel.
#pragma once

#define AV_FULL 0xffffffffu
// Pipeline stages are real functions, not inlined into the kernel: each runs once per substep (call overhead is noise)
// and the kernel's instruction footprint -- what ncu shows as the dominant stall -- stays the sum of the stages instead
// of growing with every call site; it also makes cuobjdump / ncu attribute code per stage.
#ifndef AV_STAGE
#define AV_STAGE __noinline__
#endif
#define AV_MINVAL 1e-15f

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
__device__ __forceinline__ V3 ld3(const float *p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float *p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 normalized(V3 a) {
    float n = norm(a);
    return n > AV_MINVAL ? a * (1.0f / n) : a;
}
__device__ __forceinline__ float comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

// row-major 3x3
struct M3 {
    float m[9];
};
__device__ __forceinline__ M3 ldm3(const float *p) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.m[i] = p[i];
    return r;
}
__device__ __forceinline__ void stm3(float *p, const M3 &a) {
#pragma unroll
    for (int i = 0; i < 9; i++) p[i] = a.m[i];
}
__device__ __forceinline__ V3 mul(const M3 &a, V3 v) {
    return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
              a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
__device__ __forceinline__ V3 mulT(const M3 &a, V3 v) {
    return v3(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
              a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z);
}
__device__ __forceinline__ M3 mul(const M3 &a, const M3 &b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
    return r;
}
__device__ __forceinline__ V3 colm(const M3 &a, int k) { return v3(a.m[k], a.m[3 + k], a.m[6 + k]); }

struct Q4 {
    float w, x, y, z;
};
__device__ __forceinline__ Q4 ldq(const float *p) { Q4 q = {p[0], p[1], p[2], p[3]}; return q; }
__device__ __forceinline__ void stq(float *p, Q4 q) { p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; }
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
    Q4 r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return r;
}
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
    float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    if (n < AV_MINVAL) { Q4 i = {1, 0, 0, 0}; return i; }
    float s = 1.0f / n;
    Q4 r = {q.w * s, q.x * s, q.y * s, q.z * s};
    return r;
}
__device__ __forceinline__ M3 q2m(Q4 q) {
    M3 r;
    float w = q.w, x = q.x, y = q.y, z = q.z;
    r.m[0] = 1 - 2 * (y * y + z * z); r.m[1] = 2 * (x * y - w * z); r.m[2] = 2 * (x * z + w * y);
    r.m[3] = 2 * (x * y + w * z); r.m[4] = 1 - 2 * (x * x + z * z); r.m[5] = 2 * (y * z - w * x);
    r.m[6] = 2 * (x * z - w * y); r.m[7] = 2 * (y * z + w * x); r.m[8] = 1 - 2 * (x * x + y * y);
    return r;
}

// 6-vectors: [angular; linear]
struct S6 {
    V3 a, l;
};
__device__ __forceinline__ S6 ld6(const float *p) { S6 r = {ld3(p), ld3(p + 3)}; return r; }
__device__ __forceinline__ void st6(float *p, S6 s) { st3(p, s.a); st3(p + 3, s.l); }
__device__ __forceinline__ S6 operator+(S6 a, S6 b) { S6 r = {a.a + b.a, a.l + b.l}; return r; }
__device__ __forceinline__ S6 operator*(S6 a, float s) { S6 r = {a.a * s, a.l * s}; return r; }
__device__ __forceinline__ float dot6(S6 a, S6 b) { return dot(a.a, b.a) + dot(a.l, b.l); }
__device__ __forceinline__ S6 cross_motion(S6 v, S6 s) {
    S6 r = {cross(v.a, s.a), cross(v.a, s.l) + cross(v.l, s.a)};
    return r;
}
__device__ __forceinline__ S6 cross_force(S6 v, S6 f) {
    S6 r = {cross(v.a, f.a) + cross(v.l, f.l), cross(v.a, f.l)};
    return r;
}
// spatial inertia {Ixx Iyy Izz Ixy Ixz Iyz, m*c (3), m} about the tree origin, times a motion vector
__device__ __forceinline__ S6 inert_mul(const float *I, S6 v) {
    V3 mc = v3(I[6], I[7], I[8]);
    S6 f;
    f.a = v3(I[0] * v.a.x + I[3] * v.a.y + I[4] * v.a.z, I[3] * v.a.x + I[1] * v.a.y + I[5] * v.a.z,
             I[4] * v.a.x + I[5] * v.a.y + I[2] * v.a.z) + cross(mc, v.l);
    f.l = v.l * I[9] - cross(mc, v.a);
    return f;
}

// read-only model data (hull vertices): non-coherent path.  AV_HULL_HINT=1 adds the L1 evict_last priority: the hulls are
// shared by all environments of an SM, the per-environment contact blocks that stream through the same L1 are not.
#ifndef AV_HULL_HINT
#define AV_HULL_HINT 0
#endif
__device__ __forceinline__ float4 ldg4(const float4 *p) {
#ifdef __CUDA_ARCH__
#if AV_HULL_HINT
    float4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
#else
    return *p;
#endif
}
// loads of the solver's contact blocks inside the sweeps (per-environment scratch, re-read every sweep).  AV_CBLK_LD:
// 0 default (allocate in L1), 1 ld.cg (L2 only), 2 L1::no_allocate, 3 L1::evict_first
#ifndef AV_CBLK_LD
#define AV_CBLK_LD 0
#endif
__device__ __forceinline__ float4 ldblk4(const float4 *p) {
#if defined(__CUDA_ARCH__) && AV_CBLK_LD == 1
    return __ldcg(p);
#elif defined(__CUDA_ARCH__) && AV_CBLK_LD == 2
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#elif defined(__CUDA_ARCH__) && AV_CBLK_LD == 3
    float4 v;
    asm volatile("ld.global.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ float ldblk1(const float *p) {
#if defined(__CUDA_ARCH__) && AV_CBLK_LD == 1
    return __ldcg(p);
#elif defined(__CUDA_ARCH__) && AV_CBLK_LD == 2
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#elif defined(__CUDA_ARCH__) && AV_CBLK_LD == 3
    float v;
    asm volatile("ld.global.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}

// ---- bulk asynchronous copy (TMA engine, 1-D) of one contact block from the L2-resident scratch into shared memory,
// completion signalled on an mbarrier.  Used by the solver sweep to prefetch the next contact's block while the
// current one is being updated.  The host emulation harness replaces these with plain copies.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy accesses (this warp's earlier loads / stores) ordered before the async proxy touches the same memory
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// bounded wait: returns false if the phase did not complete (the caller then reads the block from global memory)
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, unsigned parity) {
    for (int spin = 0; spin < 4096; spin++) {
        unsigned ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (ok) return true;
    }
    return false;
}
// shared -> global bulk copy (bulk async-group completion) and the wait for its source reads
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
#else
static inline void bulk_s2g(void *dst, const void *src, unsigned bytes) { memcpy(dst, src, bytes); }
static inline void bulk_commit_wait() {}
static inline void mbar_init(unsigned long long *, int) {}
static inline void fence_proxy_async() {}
static inline void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *) { memcpy(dst, src, bytes); }
static inline bool mbar_wait(unsigned long long *, unsigned) { return true; }
#endif

// ---- warp collectives (one warp == one environment)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(AV_FULL, v, o);
    return v;
}
// argmax with lowest-index tie break; returns the winning (value, index) in every lane
// Same result (largest value, smallest index among ties) with two REDUX instructions instead of ten dependent shuffles: the
// float is mapped to an order-preserving unsigned key (v must not be -0.0: callers add +0.0f).  Indices must be >= 0.
__device__ __forceinline__ void warp_argmax_redux(float &v, int &i) {
    unsigned u = __float_as_uint(v);
    u ^= (unsigned)((int)u >> 31) | 0x80000000u;
    unsigned m = __reduce_max_sync(AV_FULL, u);
    i = (int)__reduce_min_sync(AV_FULL, u == m ? (unsigned)i : 0xffffffffu);
    m ^= (m & 0x80000000u) ? 0x80000000u : 0xffffffffu;
    v = __uint_as_float(m);
}
__device__ __forceinline__ void warp_argmax(float &v, int &i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(AV_FULL, v, o);
        int oi = __shfl_xor_sync(AV_FULL, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}
