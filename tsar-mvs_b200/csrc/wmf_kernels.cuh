// wmf_kernels.cuh -- weighted-median consistency filter / fill (gipuma_WMF, gipuma_WMF_Final;
// gipuma.cu:1500-1698, 1295-1497; launch sites gipuma.cu:1809-1812, 1844-1847 -- dormant in the reference).
//
// Per pixel the reference gathers up to 11x11 "reliable" neighbours on a coarse grid into nine 144-float local
// arrays and bubble-sorts four of them.  Two properties of that code define the result and are reproduced:
//   * the sort loops run `j < num - i` and touch element [num] (the zero-initialised slot after the last
//     neighbour): effectively num+1 elements are sorted -- a dummy (value 0, weight 0, index 0)
//     takes part, and the LARGEST element ends at position num, outside every later `i < num` loop;
//   * the sort is stable (`>` comparison), sums run in sorted order.
// Instead of sorting (O(n^2) swaps of local-memory arrays, four times) the sorted sequence is generated on
// the fly by repeated stable min-extraction with a 128-bit "used" mask, and only as far as each weighted
// median needs; the only full pass is the depth order (its weights feed wSum in sorted order).
// The reference also races here (a launch reads scale/depth/norm4 of neighbours while other threads rewrite
// them); as for the propagation kernel we read a pre-launch snapshot (the caller passes *_in copies).
#pragma once
#include "glue_kernels.cuh"

namespace tsar {

constexpr int kWmfMax = 122;  // 11 x 11 neighbours + the dummy slot

struct WmfList {
    float w[kWmfMax], d[kWmfMax], x[kWmfMax], y[kWmfMax], z[kWmfMax];
    int idx[kWmfMax];
    int num;
};

// position-ordered stable extraction: returns the element with the smallest key among the unused ones
// (ties: lowest original index), marks it used.  m = num + 1 elements.
__device__ __forceinline__ int wmf_extract_min(const float *key, int m, unsigned used[4]) {
    int best = -1;
    float bk = 0.f;
    for (int e = 0; e < m; e++) {
        if (used[e >> 5] & (1u << (e & 31))) continue;
        const float k = key[e];
        if (best < 0 || k < bk) { best = e; bk = k; }
    }
    used[best >> 5] |= 1u << (best & 31);
    return best;
}

// first sorted position i < num with cumulative weight >= half; returns the element there or -1
__device__ __forceinline__ int wmf_weighted_median(const float *key, const float *w, int num, float half) {
    unsigned used[4] = {0, 0, 0, 0};
    float acc = 0.f;
    for (int i = 0; i < num; i++) {
        const int e = wmf_extract_min(key, num + 1, used);
        acc = fadd(acc, w[e]);
        if (acc >= half) return e;
    }
    return -1;
}

__device__ __forceinline__ float wmf_weight(float ref_pix, float cen_pix, int i, int j, float scale_div) {
    // exp(-spatial / (2*2)) * exp(-|ref - cen| / (3*3)),  spatial = sqrtf(i*i + j*j) / scale_div
    const float spatial = fdiv(__fsqrt_rn((float)(i * i + j * j)), scale_div);
    return fmul(expf(fmul(spatial, -0.25f)), expf(-fdiv(fabsf(fsub(ref_pix, cen_pix)), 9.0f)));
}

// gathers the neighbour list exactly in the reference's order (i = x offset outer, j = y offset inner)
__device__ __forceinline__ void wmf_gather(const GlueConst &g, const float *__restrict__ ref, const float4 *__restrict__ plane,
                                           const float *__restrict__ depth, const float *__restrict__ scale, int px, int py,
                                           int radius, int gap, float scale_div, WmfList &L) {
    const float cen = ref[(size_t)py * g.W + px];
    int num = 0;
    for (int i = -radius; i <= radius; i += gap)
        for (int j = -radius; j <= radius; j += gap) {
            const int nx = px + i, ny = py + j;
            if (nx < 0 || nx >= g.W || ny < 0 || ny >= g.H) continue;
            const int ne = ny * g.W + nx;
            if (scale[ne] != 1.0f) continue;
            L.w[num] = wmf_weight(ref[ne], cen, i, j, scale_div);
            L.d[num] = depth[ne];
            L.idx[num] = ne;
            const float4 n = plane[ne];
            L.x[num] = n.x; L.y[num] = n.y; L.z[num] = n.z;
            num++;
        }
    L.w[num] = 0.f; L.d[num] = 0.f; L.idx[num] = 0; L.x[num] = 0.f; L.y[num] = 0.f; L.z[num] = 0.f;  // slot [num] = {0}
    L.num = num;
}

// the common middle part of both kernels: weighted medians of the normal components and of the depth.
// Returns false when the depth median is never reached (the reference then leaves norm_mid.w unset).
__device__ __forceinline__ bool wmf_median_plane(const GlueConst &g, const float *__restrict__ depth, const WmfList &L,
                                                 float4 &norm_mid) {
    const int num = L.num;
    // wSum: weights in ascending-depth order, positions 0..num-1 (the largest depth is excluded)
    unsigned used[4] = {0, 0, 0, 0};
    unsigned char order[kWmfMax];
    float wsum = 0.f;
    for (int i = 0; i <= num; i++) {
        const int e = wmf_extract_min(L.d, num + 1, used);
        order[i] = (unsigned char)e;
        if (i < num) wsum = fadd(wsum, L.w[e]);
    }
    const float half = fmul(wsum, 0.5f);
    int e;
    norm_mid = make_float4(0.f, 0.f, 0.f, 0.f);  // (uninitialised in the reference when a median is never reached)
    if ((e = wmf_weighted_median(L.x, L.w, num, half)) >= 0) norm_mid.x = L.x[e];
    if ((e = wmf_weighted_median(L.y, L.w, num, half)) >= 0) norm_mid.y = L.y[e];
    if ((e = wmf_weighted_median(L.z, L.w, num, half)) >= 0) norm_mid.z = L.z[e];
    float acc = 0.f;
    for (int i = 0; i < num; i++) {
        acc = fadd(acc, L.w[order[i]]);
        if (acc >= half) {
            const int weimid = L.idx[order[i]];
            const float disp_mid = fdiv(fmul(g.f_params, g.baseline), depth[weimid]);
            const int mx = weimid % g.W, my = weimid / g.W;
            const float len = __fsqrt_rn(dot3(norm_mid.x, norm_mid.x, norm_mid.y, norm_mid.y, norm_mid.z, norm_mid.z));
            norm_mid.x = fdiv(norm_mid.x, len); norm_mid.y = fdiv(norm_mid.y, len); norm_mid.z = fdiv(norm_mid.z, len);
            norm_mid.w = g_plane_d(g, norm_mid.x, norm_mid.y, norm_mid.z, mx, my, disp_mid);
            return true;
        }
    }
    return false;
}

// gipuma_WMF (gipuma.cu:1500-1698): reliability test of every pixel against the weighted-median plane of its
// reliable neighbours on a coarse grid; writes scale only.
__global__ void __launch_bounds__(64) wmf_kernel(const __grid_constant__ GlueConst g, const float *__restrict__ ref,
                                                 const float4 *__restrict__ plane, const float *__restrict__ depth,
                                                 const float *__restrict__ scale_in, float *__restrict__ scale_out, int iter) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const int p = y * g.W + x;
    const int po = 1 << iter, repo = 1 << (3 - iter);
    WmfList L;
    wmf_gather(g, ref, plane, depth, scale_in, x, y, 80 / po, 16 / po, (float)repo, L);
    float4 norm_mid;
    wmf_median_plane(g, depth, L, norm_mid);
    const int ths = 24 / po;
    float out = 0.f;
    if (L.num > 0) {
        const float fb = fmul(g.f_params, g.baseline);
        const float disp_now = fdiv(fb, g_plane_depth(g, norm_mid, x, y));
        const float disp_org = fdiv(fb, g_plane_depth(g, plane[p], x, y));
        out = (fabsf(fsub(disp_now, disp_org)) > (float)ths) ? 0.f : 1.f;
    }
    scale_out[p] = out;
}

// gipuma_WMF_Final (gipuma.cu:1295-1497): fills unreliable pixels of textured regions from the weighted median
// of reliable neighbours, fine to coarse.
__global__ void __launch_bounds__(64) wmf_final_kernel(const __grid_constant__ GlueConst g, const float *__restrict__ ref,
                                                       const float4 *__restrict__ plane_in, const float *__restrict__ depth_in,
                                                       const float *__restrict__ scale_in, float4 *__restrict__ plane_out,
                                                       float *__restrict__ depth_out, float *__restrict__ scale_out,
                                                       const float *__restrict__ canny, const float *__restrict__ region_text,
                                                       int n_regions, int iter) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const int p = y * g.W + x;
    const int reg = (int)canny[p];
    if (reg < 0 || reg >= n_regions || region_text[reg] != 1.0f || scale_in[p] != 0.0f) return;
    const int po = 1 << iter;
    WmfList L;
    wmf_gather(g, ref, plane_in, depth_in, scale_in, x, y, 5 * po, po, (float)po, L);
    if (L.num < 32 / po) return;
    float4 norm_mid;
    if (!wmf_median_plane(g, depth_in, L, norm_mid)) return;
    plane_out[p] = norm_mid;
    float d = fdiv(fmul(g.f_params, g.baseline), g_plane_depth(g, norm_mid, x, y));
    float sc = 1.f;
    if (d <= g.min_disp || d >= g.max_disp) { sc = 0.f; d = g.min_disp; }
    depth_out[p] = d;
    scale_out[p] = sc;
}

// snapshot = pre-launch copies of the arrays the launch rewrites (scratch supplied by the context)
static inline int wmf_launch(const GlueConst &g, const float *ref, float4 *plane, float *depth, float *scale, float *scale_snapshot,
                             int iter, cudaStream_t s) {
    const size_t n = (size_t)g.W * g.H;
    if (cudaMemcpyAsync(scale_snapshot, scale, n * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return TSAR_ERR_CUDA;
    dim3 b(32, 2), grid((g.W + 31) / 32, (g.H + 1) / 2);
    wmf_kernel<<<grid, b, 0, s>>>(g, ref, plane, depth, scale_snapshot, scale, iter);
    return cudaGetLastError() == cudaSuccess ? TSAR_OK : TSAR_ERR_CUDA;
}

static inline int wmf_final_launch(const GlueConst &g, const float *ref, float4 *plane, float *depth, float *scale,
                                   float4 *plane_snapshot, float *depth_snapshot, float *scale_snapshot, const float *canny,
                                   const float *region_text, int n_regions, int iter, cudaStream_t s) {
    const size_t n = (size_t)g.W * g.H;
    if (cudaMemcpyAsync(plane_snapshot, plane, n * 16, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(depth_snapshot, depth, n * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(scale_snapshot, scale, n * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
        return TSAR_ERR_CUDA;
    dim3 b(32, 2), grid((g.W + 31) / 32, (g.H + 1) / 2);
    wmf_final_kernel<<<grid, b, 0, s>>>(g, ref, plane_snapshot, depth_snapshot, scale_snapshot, plane, depth, scale, canny,
                                        region_text, n_regions, iter);
    return cudaGetLastError() == cudaSuccess ? TSAR_OK : TSAR_ERR_CUDA;
}

}  // namespace tsar
