// avsim_obs.cuh -- device-resident observation path.
//
// Replaces the image half of lerobot's preprocess_observation (reference lerobot/lerobot/common/envs/utils.py:37-50):
// torch.from_numpy(img) -> rearrange "b h w c -> b c h w" -> .type(float32) -> /= 255, done on the host for every
// camera of every environment of every step (3.7 MB of uint8 in, 14.7 MB of fp32 out per environment at 4 cameras).
// Here the uint8 frames written by avsim_render stay in HBM and one kernel produces the policy's input planes.
//
// Bound: HBM.  Algorithmic bytes per image = 3*H*W read + 12*H*W written (480x640: 0.92 MB + 3.69 MB).
// Layout: a thread owns 4 consecutive pixels = three aligned 32-bit words of the interleaved source (a warp reads 384
// contiguous bytes per load instruction) and writes one float4 to each colour plane (512 contiguous bytes per warp and
// plane).  AV_OBS_UNROLL independent quads per thread keep 12 loads in flight before the first store.  Loads and stores
// are streaming (.cs): every byte is touched exactly once.
//
// Arithmetic is bit-exact with the reference's fp32 division: u8 -> f32 by exponent splice (0x4B000000 | v) - 2^23
// (exact), then q = v * (1/255) with one Markstein correction  q += fma(-255, q, v) * (1/255), which is the correctly
// rounded quotient for all 256 inputs (checked exhaustively in tests/test_obs_path.py, and against torch on the GPU).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define AV_OBS_HD __host__ __device__ __forceinline__
#else
#define AV_OBS_HD static inline
#endif

#define AV_OBS_UNROLL 4
#define AV_OBS_THREADS 256

AV_OBS_HD float av_u8_unit(float v) {   // v / 255, correctly rounded, v an integer in [0, 255]
    const float r = 1.0f / 255.0f;                                 // rounded reciprocal (compile-time constant)
#ifdef __CUDA_ARCH__
    float q = v * r;
    return fmaf(fmaf(-255.0f, q, v), r, q);
#else
    float q = v * r;
    return __builtin_fmaf(__builtin_fmaf(-255.0f, q, v), r, q);
#endif
}

#ifdef __CUDACC__   // the kernels; the arithmetic above also compiles with g++ for the exhaustive CPU check
__device__ __forceinline__ float av_byte_f32(uint32_t w, uint32_t sel) {   // byte `sel & 3` of w as an exact float
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - 8388608.0f;
}

// src u8 [n][H][W][3] -> dst f32 [n][3][H][W]; hw4 = H*W/4 (H*W a multiple of 4), nquads = n * hw4
__global__ void __launch_bounds__(AV_OBS_THREADS) avsim_pixels_to_float_kernel(const uint32_t *__restrict__ src, float *__restrict__ dst,
                                                                               long long nquads, int hw4) {
    const long long stride = (long long)gridDim.x * AV_OBS_THREADS;
    for (long long q0 = (long long)blockIdx.x * AV_OBS_THREADS + threadIdx.x; q0 < nquads; q0 += stride * AV_OBS_UNROLL) {
        uint32_t w[AV_OBS_UNROLL][3];
#pragma unroll
        for (int u = 0; u < AV_OBS_UNROLL; u++) {
            long long q = q0 + u * stride;
            if (q < nquads) {
                w[u][0] = __ldcs(src + 3 * q);
                w[u][1] = __ldcs(src + 3 * q + 1);
                w[u][2] = __ldcs(src + 3 * q + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < AV_OBS_UNROLL; u++) {
            long long q = q0 + u * stride;
            if (q >= nquads) break;
            long long img = q / hw4;
            int r = (int)(q - img * hw4);
            float4 *o = reinterpret_cast<float4 *>(dst + img * 12 * hw4) + r;     // plane stride = hw4 float4s
            const uint32_t a = w[u][0], b = w[u][1], c = w[u][2];                  // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
            float4 R, G, Bl;
            R.x = av_u8_unit(av_byte_f32(a, 0x7440)); R.y = av_u8_unit(av_byte_f32(a, 0x7443));
            R.z = av_u8_unit(av_byte_f32(b, 0x7442)); R.w = av_u8_unit(av_byte_f32(c, 0x7441));
            G.x = av_u8_unit(av_byte_f32(a, 0x7441)); G.y = av_u8_unit(av_byte_f32(b, 0x7440));
            G.z = av_u8_unit(av_byte_f32(b, 0x7443)); G.w = av_u8_unit(av_byte_f32(c, 0x7442));
            Bl.x = av_u8_unit(av_byte_f32(a, 0x7442)); Bl.y = av_u8_unit(av_byte_f32(b, 0x7441));
            Bl.z = av_u8_unit(av_byte_f32(c, 0x7440)); Bl.w = av_u8_unit(av_byte_f32(c, 0x7443));
            __stcs(o, R);
            __stcs(o + hw4, G);
            __stcs(o + 2 * hw4, Bl);
        }
    }
}

// any H*W (e.g. the 225x300 eval frame, reference env.py:195-200): thread per pixel-channel of the OUTPUT (coalesced
// stores, byte gathers served by L1)
__global__ void __launch_bounds__(AV_OBS_THREADS) avsim_pixels_to_float_any_kernel(const uint8_t *__restrict__ src, float *__restrict__ dst,
                                                                                   long long nvals, int hw) {
    const long long stride = (long long)gridDim.x * AV_OBS_THREADS;
    for (long long i = (long long)blockIdx.x * AV_OBS_THREADS + threadIdx.x; i < nvals; i += stride) {
        long long img = i / (3ll * hw);
        int rem = (int)(i - img * 3ll * hw), c = rem / hw, p = rem - c * hw;
        dst[i] = av_u8_unit((float)src[img * 3ll * hw + 3ll * p + c]);
    }
}
#endif  // __CUDACC__
