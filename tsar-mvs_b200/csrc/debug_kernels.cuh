// debug_kernels.cuh -- instrumentation: raw texture sampling (to calibrate a software model of the
// texture unit's bilinear filter) and issue-rate microbenchmarks (FP32 FFMA, MUFU, TEX) that give the
// measured roofline denominators for the PatchMatch kernels (MEASURED_PEAKS.json only has HBM / bf16).
#pragma once
#include <cuda_runtime.h>

namespace tsar {

__global__ void dbg_tex_sample_kernel(cudaTextureObject_t tex, int n, const float2 *__restrict__ xy, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex2D<float>(tex, xy[i].x, xy[i].y);
}

// 8 independent FFMA chains per thread, `iters` rounds: 16 flops per round per thread
__global__ void __launch_bounds__(256) dbg_ffma_kernel(float *out, int iters, float a, float b) {
    float r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll 8
        for (int k = 0; k < 8; k++) {
            r0 = fmaf(r0, a, b); r1 = fmaf(r1, a, b); r2 = fmaf(r2, a, b); r3 = fmaf(r3, a, b);
            r4 = fmaf(r4, a, b); r5 = fmaf(r5, a, b); r6 = fmaf(r6, a, b); r7 = fmaf(r7, a, b);
        }
    }
    if (r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7 == 12345.678f) out[0] = r0;
}

// 8 independent MUFU.RSQ chains per thread (rsqrt is not an involution, so nothing can be folded away -- the first
// version chained rcp(rcp(x)) and the compiler removed the whole loop): 8 x 8 = 64 MUFU per round per thread
__device__ __forceinline__ float dbg_rsq(float v) {
    float r;
    asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__global__ void __launch_bounds__(256) dbg_mufu_kernel(float *out, int iters) {
    float r0 = 1.0f + threadIdx.x * 1e-3f, r1 = r0 + .1f, r2 = r0 + .2f, r3 = r0 + .3f, r4 = r0 + .4f, r5 = r0 + .5f, r6 = r0 + .6f,
          r7 = r0 + .7f;
    for (int i = 0; i < iters; i++) {
#pragma unroll 8
        for (int k = 0; k < 8; k++) {
            r0 = dbg_rsq(r0); r1 = dbg_rsq(r1); r2 = dbg_rsq(r2); r3 = dbg_rsq(r3);
            r4 = dbg_rsq(r4); r5 = dbg_rsq(r5); r6 = dbg_rsq(r6); r7 = dbg_rsq(r7);
        }
    }
    if (r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7 == 12345.678f) out[0] = r0;
}

// bilinear fp32 fetches with warp-local coordinates (like a warped window): `iters` x 4 TEX per thread
__global__ void __launch_bounds__(256) dbg_tex_kernel(cudaTextureObject_t tex, float *out, int iters, int W, int H) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float x = (float)(t % W) + 0.37f, y = (float)((t / W) % H) + 0.61f;
    float acc = 0.f;
    for (int i = 0; i < iters; i++) {
        acc += tex2D<float>(tex, x, y);
        acc += tex2D<float>(tex, x + 2.13f, y + 0.21f);
        acc += tex2D<float>(tex, x + 0.17f, y + 2.29f);
        acc += tex2D<float>(tex, x + 2.41f, y + 2.07f);
        x += 0.93f; y += 0.11f;
        if (x > W) x -= W;
    }
    if (acc == 12345.678f) out[0] = acc;
}

}  // namespace tsar
