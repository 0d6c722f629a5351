// ransac_kernels.cuh -- per-region plane fitting for textureless regions ("per-segment plane fitting",
// north-star item 4).  Reference: the CPU loop in main.cpp:1520-1730
// (calcLinePara main.cpp:147-164): for every region flagged textureless, back-project its reliable pixels,
// run 10 000 RANSAC triples with an adaptive inlier threshold, then 1000 x 4 local perturbation rounds;
// about 7e8 double residuals per region on one CPU core in the reference.
//
// Here: device-side stable compaction of the region's reliable pixels (raster order, as the reference collects
// them), back-projection, and one persistent 1024-thread CTA per region that evaluates every hypothesis with a
// block-wide inlier count.  All arithmetic is IEEE double without contraction, in the reference's evaluation
// order, so the result is bit-identical to a scalar C restatement given the same random numbers.  The reference
// draws them from rand()/system_clock (not reproducible); the caller supplies the stream instead (the values
// rand() would have returned, 46 000 per region), which also makes the fit testable.
// Kept quirks: the A coefficient of calcLinePara uses (y3-y1) twice (main.cpp:159); ties (>=) replace the best.
// Deviation: when a region has more than 49 999 reliable pixels the reference keeps a random subset
// (std::shuffle with a clock seed); we keep an evenly strided subset.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "glue_kernels.cuh"

namespace tsar {

constexpr int kRansacMaxPts = 49999;       // main.cpp:1541-1549: lists are cut to fewer than 50 000 points
constexpr int kRansacIters = 10000;        // main.cpp:1603
constexpr int kRefineRounds = 1000;        // main.cpp:1662
constexpr int kRansacRandPerRegion = 3 * kRansacIters + 4 * 4 * kRefineRounds;

// ---- stable compaction of one region's reliable pixels ------------------------------------------------------
__global__ void ransac_flag_kernel(const float *__restrict__ scale, const float *__restrict__ canny, int n, int region,
                                   int *__restrict__ block_counts) {
    __shared__ int warp_sums[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = (i < n && scale[i] == 1.0f && (int)canny[i] == region) ? 1 : 0;
    const unsigned b = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += warp_sums[w];
        block_counts[blockIdx.x] = s;
    }
}

__global__ void ransac_scan_blocks_kernel(int *block_counts, int nblocks, int *total) {  // single thread block
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int v = i < nblocks ? block_counts[i] : 0;
        // inclusive scan inside the block (Hillis-Steele through shared memory)
        __shared__ int buf[1024];
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < (int)blockDim.x; o <<= 1) {
            const int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblocks) block_counts[i] = carry + buf[threadIdx.x] - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += buf[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void ransac_scatter_kernel(const float *__restrict__ scale, const float *__restrict__ canny, int n, int region,
                                      const int *__restrict__ block_offsets, int *__restrict__ list) {
    __shared__ int warp_off[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = (i < n && scale[i] == 1.0f && (int)canny[i] == region) ? 1 : 0;
    const unsigned b = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_off[warp] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { const int t = warp_off[w]; warp_off[w] = s; s += t; }
    }
    __syncthreads();
    if (f) list[block_offsets[blockIdx.x] + warp_off[warp] + __popc(b & ((1u << lane) - 1))] = i;
}

// back-projection of the (sub-sampled) list: main.cpp:1574-1598, float arithmetic, no contraction
__global__ void ransac_points_kernel(const __grid_constant__ GlueConst g, const float *__restrict__ depth,
                                     const int *__restrict__ list, int n_all, int n_used, float3 *__restrict__ pts) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_used) return;
    const int src = (n_all == n_used) ? k : (int)(((long long)k * n_all) / n_used);
    const int p = list[src];
    const float disp = fdiv(fmul(g.f_params, g.baseline), depth[p]);
    const int px = p % g.W, py = p / g.W;
    const float x = fsub(fmul(disp, (float)px), g.Pc[0]), y = fsub(fmul(disp, (float)py), g.Pc[1]), z = fsub(disp, g.Pc[2]);
    float3 o;
    o.x = fadd(fadd(fmul(g.Minv[0], x), fmul(g.Minv[1], y)), fmul(g.Minv[2], z));
    o.y = fadd(fadd(fmul(g.Minv[3], x), fmul(g.Minv[4], y)), fmul(g.Minv[5], z));
    o.z = fadd(fadd(fmul(g.Minv[6], x), fmul(g.Minv[7], y)), fmul(g.Minv[8], z));
    pts[k] = o;
}

// ---- the fit: one CTA per region -----------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// number of points with |x*a + y*b + z*c + d| < thr   (block-wide, returned to every thread)
__device__ __forceinline__ int ransac_count(const float3 *__restrict__ pts, int n, double a, double b, double c, double d,
                                            double thr, int *smem_counts) {
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float3 p = pts[i];
        const double r = fabs(dadd(dadd(dadd(dmul((double)p.x, a), dmul((double)p.y, b)), dmul((double)p.z, c)), d));
        cnt += (r < thr) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem_counts[threadIdx.x >> 5] = cnt;
    __syncthreads();
    int tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += smem_counts[w];
    return tot;
}

struct RansacJob {
    const float3 *pts;   // region's points
    int n;               // number of points (0: region skipped, plane left unchanged)
    float size;          // cannylines->size[region]
    const uint32_t *rnd; // kRansacRandPerRegion values, in the order the reference calls rand()
    float4 *out;         // cannylines->norm4[region]
};

__global__ void __launch_bounds__(1024) ransac_fit_kernel(const RansacJob *__restrict__ jobs) {
    __shared__ int smem_counts[32];
    const RansacJob job = jobs[blockIdx.x];
    const int n = job.n;
    if (n <= 0) return;
    const float3 *pts = job.pts;
    const uint32_t *rnd = job.rnd;
    float depth_abs = (float)(0.0003 * (double)sqrtf(job.size / 20.0f));  // main.cpp:1552-1553
    double a = 0, b = 0, c = 1, d = -1, maximum = 0;
    for (int k = 0; k < kRansacIters; k++) {
        const float3 P1 = pts[rnd[3 * k] % (uint32_t)n], P2 = pts[rnd[3 * k + 1] % (uint32_t)n], P3 = pts[rnd[3 * k + 2] % (uint32_t)n];
        const double x1 = P1.x, y1 = P1.y, z1 = P1.z, x2 = P2.x, y2 = P2.y, z2 = P2.z, x3 = P3.x, y3 = P3.y, z3 = P3.z;
        // calcLinePara (main.cpp:147-164), including its A coefficient as written
        double ta = dadd(dmul(dadd(y3, -y1), dadd(z3, -z1)), -dmul(dadd(z2, -z1), dadd(y3, -y1)));
        double tb = dadd(dmul(dadd(x3, -x1), dadd(z2, -z1)), -dmul(dadd(x2, -x1), dadd(z3, -z1)));
        double tc = dadd(dmul(dadd(x2, -x1), dadd(y3, -y1)), -dmul(dadd(x3, -x1), dadd(y2, -y1)));
        double td = -dadd(dadd(dmul(ta, x1), dmul(tb, y1)), dmul(tc, z1));
        const double sq = __dsqrt_rn(dadd(dadd(dmul(ta, ta), dmul(tb, tb)), dmul(tc, tc)));
        ta = __ddiv_rn(ta, sq); tb = __ddiv_rn(tb, sq); tc = __ddiv_rn(tc, sq); td = __ddiv_rn(td, sq);
        const double cnt = (double)ransac_count(pts, n, ta, tb, tc, td, (double)depth_abs, smem_counts);
        if (cnt >= maximum) { a = ta; b = tb; c = tc; d = td; maximum = cnt; }
        if (k % 1000 == 0) {  // adaptive threshold, main.cpp:1645-1663
            const double rat = maximum / (double)n;
            if (rat < 0.3 && (double)depth_abs < 0.003) {
                depth_abs = (float)((double)depth_abs + 0.0001);
            } else {
                const double m2 = (double)ransac_count(pts, n, a, b, c, d, dadd((double)depth_abs, 0.0001), smem_counts);
                if (m2 > dadd(maximum, dmul((double)n, 0.02))) {
                    depth_abs = (float)((double)depth_abs + 0.0001);
                    maximum = m2;
                }
            }
        }
    }
    const uint32_t *rr = rnd + 3 * kRansacIters;
    for (int i = 0; i < kRefineRounds; i++)  // main.cpp:1668-1710
        for (int j = 2000; j >= 2; j /= 10) {
            const int med = j / 2;
            const double da = (double)((int)(rr[0] % (uint32_t)j) - med) / 10000.0, db = (double)((int)(rr[1] % (uint32_t)j) - med) / 10000.0;
            const double dc = (double)((int)(rr[2] % (uint32_t)j) - med) / 10000.0, dd = (double)((int)(rr[3] % (uint32_t)j) - med) / 1000.0;
            rr += 4;
            double ra = dadd(a, da), rb = dadd(b, db), rc = dadd(c, dc), rd = dadd(d, dd);
            const double sq = __dsqrt_rn(dadd(dadd(dmul(ra, ra), dmul(rb, rb)), dmul(rc, rc)));
            ra = __ddiv_rn(ra, sq); rb = __ddiv_rn(rb, sq); rc = __ddiv_rn(rc, sq); rd = __ddiv_rn(rd, sq);
            const double cnt = (double)ransac_count(pts, n, ra, rb, rc, rd, (double)depth_abs, smem_counts);
            if (cnt >= maximum) { a = ra; b = rb; c = rc; d = rd; maximum = cnt; }
        }
    if (threadIdx.x == 0) *job.out = make_float4((float)a, (float)b, (float)c, (float)d);
}

}  // namespace tsar
