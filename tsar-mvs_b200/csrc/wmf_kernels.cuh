// wmf_kernels.cuh -- weighted-median consistency filter / fill (gipuma_WMF, gipuma_WMF_Final;
// gipuma.cu:1500-1698, 1295-1497; launch sites gipuma.cu:1809-1812, 1844-1847 -- dormant in the reference).
//
// Per pixel the reference gathers up to 11x11 "reliable" neighbours on a coarse grid into nine 144-float local
// arrays and bubble-sorts four of them.  Two properties of that code define the result and are reproduced:
//   * the sort loops run `j < num - i` and touch element [num] (the zero-initialised slot after the last
//     neighbour): effectively num+1 elements are sorted -- a dummy (value 0, weight 0, index 0)
//     takes part, and the LARGEST element ends at position num, outside every later `i < num` loop;
//   * the sort is stable (`>` comparison), sums run in sorted order.
// Instead of bubble-sorting (O(n^2) swaps of local-memory arrays, four times) each key is ordered by a per-thread
// stable merge sort of (key, position) composites, O(n log n); sums then run over that order sequentially, as the
// reference's do.
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

// Stable ascending order of key[0..m-1] (ties: lower original position first), as a list of positions.
// Per-thread bottom-up merge sort of 64-bit composites (order-preserving key bits << 32 | position) between two
// local buffers: O(m log m) instead of the reference's O(m^2) bubble sort, same resulting sequence.
// (-0.0 is folded onto +0.0 so that it ties with the dummy's 0 exactly as the float comparison does.)
struct WmfOrder {
    unsigned long long a[kWmfMax], b[kWmfMax];
};

__device__ __forceinline__ const unsigned long long *wmf_sort(const float *key, int m, WmfOrder &buf) {
    for (int e = 0; e < m; e++) {
        const unsigned bits = __float_as_uint(key[e] + 0.0f);
        const unsigned ord = bits ^ ((bits >> 31) ? 0xffffffffu : 0x80000000u);
        buf.a[e] = ((unsigned long long)ord << 32) | (unsigned)e;
    }
    unsigned long long *src = buf.a, *dst = buf.b;
    for (int width = 1; width < m; width <<= 1) {
        for (int lo = 0; lo < m; lo += 2 * width) {
            const int mid = min(lo + width, m), hi = min(lo + 2 * width, m);
            int i = lo, j = mid;
            for (int o = lo; o < hi; o++) {
                const unsigned long long vi = src[min(i, mid - 1)], vj = src[min(j, m - 1)];  // clamped reads; unused when exhausted
                const bool left = (i < mid) && (j >= hi || vi <= vj);
                dst[o] = left ? vi : vj;
                i += left ? 1 : 0;
                j += left ? 0 : 1;
            }
        }
        unsigned long long *t = src; src = dst; dst = t;
    }
    return src;
}

// first sorted position i < num with cumulative weight >= half; returns the element there or -1
__device__ __forceinline__ int wmf_weighted_median(const float *key, const float *w, int num, float half, WmfOrder &buf) {
    const unsigned long long *ord = wmf_sort(key, num + 1, buf);
    float acc = 0.f;
    for (int i = 0; i < num; i++) {
        const int e = (int)(unsigned)ord[i];
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
    WmfOrder buf;
    unsigned char order[kWmfMax];
    float wsum = 0.f;
    {
        const unsigned long long *ord = wmf_sort(L.d, num + 1, buf);
        for (int i = 0; i <= num; i++) {
            const int e = (int)(unsigned)ord[i];
            order[i] = (unsigned char)e;
            if (i < num) wsum = fadd(wsum, L.w[e]);
        }
    }
    const float half = fmul(wsum, 0.5f);
    int e;
    norm_mid = make_float4(0.f, 0.f, 0.f, 0.f);  // (uninitialised in the reference when a median is never reached)
    if ((e = wmf_weighted_median(L.x, L.w, num, half, buf)) >= 0) norm_mid.x = L.x[e];
    if ((e = wmf_weighted_median(L.y, L.w, num, half, buf)) >= 0) norm_mid.y = L.y[e];
    if ((e = wmf_weighted_median(L.z, L.w, num, half, buf)) >= 0) norm_mid.z = L.z[e];
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

// ---------------------------------------------------------------------------------------------------------------
// gipuma_WMF, cooperative version.  The per-thread version above keeps ~5 KB of lists per thread in local memory;
// at any useful occupancy that is far beyond L1 (ncu: local loads 0.4 % L1 hits, 1.6 useful bytes per 32-byte
// sector, 388 GB of L2 traffic per 2-Mpx level, L2-bound).  Here a CTA of 8 warps owns 8 consecutive pixels:
//   phase 1 (thread = pixel x neighbour group): the 121 lattice neighbours of the 8 pixels are gathered into shared
//     memory, pixel-major, together with a per-pixel validity mask (42 KB per CTA -> 5 CTAs = 40 warps per SM);
//   phase 2 (warp = pixel): the (num+1)-element stable order of each key is a 128-element bitonic sort of
//     (order-preserving key bits << 32 | neighbour slot) composites held in registers, 4 per lane; the slot number
//     as low word reproduces the reference's stable order, invalid slots sort to the end, slot 121 is the dummy.
//     The sorted weight sequences go to a small per-warp scratch and ONE lane per key turns them into running
//     sums sequentially, so every floating-point sum runs in exactly the reference's order; the medians are then a
//     warp-wide search for the first running sum >= half.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kCoopPitch = 129;    // slots per pixel (121 neighbours + dummy + padding), odd -> conflict-free columns
constexpr int kCoopChain = 132;    // pitch of a sorted sequence (the 4 sequences of a warp start on distinct banks)
constexpr int kCoopSlots = 121, kCoopDummy = 121;
constexpr int kCoopRun = 8;        // pixels per CTA = warps per CTA

struct WmfCoopSmem {
    float w[kCoopRun][kCoopPitch], d[kCoopRun][kCoopPitch], x[kCoopRun][kCoopPitch], y[kCoopRun][kCoopPitch], z[kCoopRun][kCoopPitch];
    unsigned mask[kCoopRun][4];
    float acc[kCoopRun][4][kCoopChain];          // sorted weights, then their running sums (in place)
    unsigned char slot[kCoopRun][4][kCoopChain]; // neighbour slot at each sorted position
};

__device__ __forceinline__ unsigned long long wmf_composite(float key, int slot) {
    const unsigned bits = __float_as_uint(key + 0.0f);  // -0 -> +0: ties with the dummy's 0 like the float compare
    const unsigned ord = bits ^ ((bits >> 31) ? 0xffffffffu : 0x80000000u);
    return ((unsigned long long)ord << 32) | (unsigned)slot;
}

// ascending bitonic sort of 128 composites, element e = lane * 4 + r
__device__ __forceinline__ void wmf_bitonic128(unsigned long long (&v)[4], int lane) {
#pragma unroll
    for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 4) {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, v[r], j >> 2);
                    const int e = lane * 4 + r;
                    const bool up = (e & k) == 0, lower = (e & j) == 0;
                    const bool take_min = (lower == up);
                    const unsigned long long mn = v[r] < other ? v[r] : other, mx = v[r] < other ? other : v[r];
                    v[r] = take_min ? mn : mx;
                }
            } else {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    if ((r & j) == 0) {
                        const int e = lane * 4 + r;
                        const bool up = (e & k) == 0;
                        const unsigned long long a = v[r], b = v[r ^ j];
                        const unsigned long long mn = a < b ? a : b, mx = a < b ? b : a;
                        v[r] = up ? mn : mx;
                        v[r ^ j] = up ? mx : mn;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256, 5) wmf_coop_kernel(const __grid_constant__ GlueConst g, const float *__restrict__ ref,
                                                          const float4 *__restrict__ plane, const float *__restrict__ depth,
                                                          const float *__restrict__ scale_in, float *__restrict__ scale_out, int iter) {
    extern __shared__ __align__(16) unsigned char wmf_smem_raw[];
    WmfCoopSmem &sm = *reinterpret_cast<WmfCoopSmem *>(wmf_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = g.W, H = g.H;
    const int y = blockIdx.y, x0 = blockIdx.x * kCoopRun;
    const int po = 1 << iter, repo = 1 << (3 - iter);
    const int radius = 80 / po, gap = 16 / po;
    const float scale_div = (float)repo;

    // ---- phase 1: gather.  thread = (pixel tid & 7, slot group tid >> 3); a warp-wide load covers 4 neighbour
    // slots x 8 consecutive pixels = four fully used 32-byte sectors.  Slots follow the reference's list order
    // (x offset outer, y offset inner).
    if (threadIdx.x < kCoopRun * 4) sm.mask[threadIdx.x >> 2][threadIdx.x & 3] = 0u;
    if (threadIdx.x < kCoopRun) {  // the dummy element the reference's bubble sort drags in: slot [num] = {0}
        sm.w[threadIdx.x][kCoopDummy] = 0.f; sm.d[threadIdx.x][kCoopDummy] = 0.f;
        sm.x[threadIdx.x][kCoopDummy] = 0.f; sm.y[threadIdx.x][kCoopDummy] = 0.f; sm.z[threadIdx.x][kCoopDummy] = 0.f;
    }
    __syncthreads();
    {
        const int pl = threadIdx.x & (kCoopRun - 1), sg = threadIdx.x >> 3;
        const int px = x0 + pl;
        if (px < W) {
            const float cen = ref[(size_t)y * W + px];
            for (int c = sg; c < kCoopSlots; c += 32) {
                const int ii = c / 11, jj = c - ii * 11;
                const int i = -radius + ii * gap, j = -radius + jj * gap;
                const int nx = px + i, ny = y + j;
                if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
                const int ne = ny * W + nx;
                if (scale_in[ne] != 1.0f) continue;
                sm.w[pl][c] = wmf_weight(ref[ne], cen, i, j, scale_div);
                sm.d[pl][c] = depth[ne];
                const float4 n = plane[ne];
                sm.x[pl][c] = n.x; sm.y[pl][c] = n.y; sm.z[pl][c] = n.z;
                atomicOr(&sm.mask[pl][c >> 5], 1u << (c & 31));
            }
        }
    }
    __syncthreads();

    // ---- phase 2: one warp per pixel
    const int l = warp;
    const int px = x0 + l;
    if (px >= W) return;
    const unsigned m0 = sm.mask[l][0], m1 = sm.mask[l][1], m2 = sm.mask[l][2], m3 = sm.mask[l][3];
    const int num = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
    const int p = y * W + px;
    if (num == 0) {
        if (lane == 0) scale_out[p] = 0.f;
        return;
    }
    // validity of this lane's four slots (slot = lane*4 + r lies in mask word lane/8)
    const unsigned mw = (lane < 8) ? m0 : (lane < 16) ? m1 : (lane < 24) ? m2 : m3;
    const unsigned vbits = (mw >> ((lane & 7) * 4)) & 0xFu;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float(*key)[kCoopPitch] = q == 0 ? sm.d : (q == 1 ? sm.x : (q == 2 ? sm.y : sm.z));
        unsigned long long v[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int slot = lane * 4 + r;
            if ((vbits >> r) & 1u) v[r] = wmf_composite(key[l][slot], slot);
            else if (slot == kCoopDummy) v[r] = wmf_composite(0.f, slot);
            else v[r] = 0xffffffffffffff00ull | (unsigned)slot;
        }
        wmf_bitonic128(v, lane);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int s_ = lane * 4 + r;
            if (s_ < num) {  // positions 0..num-1 are the ones every later loop reads (the largest element is left out)
                const int slot = (int)(v[r] & 0xffu);
                sm.acc[l][q][s_] = sm.w[l][slot];
                sm.slot[l][q][s_] = (unsigned char)slot;
            }
        }
    }
    __syncwarp();
    // running sums in sorted order, sequentially (the reference's order of additions): lane q owns sequence q
    if (lane < 4) {
        float *seq = sm.acc[l][lane];
        float acc = 0.f;
        int s_ = 0;
        for (; s_ + 4 <= num; s_ += 4) {
            const float w0 = seq[s_], w1 = seq[s_ + 1], w2 = seq[s_ + 2], w3 = seq[s_ + 3];
            acc = fadd(acc, w0); seq[s_] = acc;
            acc = fadd(acc, w1); seq[s_ + 1] = acc;
            acc = fadd(acc, w2); seq[s_ + 2] = acc;
            acc = fadd(acc, w3); seq[s_ + 3] = acc;
        }
        for (; s_ < num; s_++) { acc = fadd(acc, seq[s_]); seq[s_] = acc; }
    }
    __syncwarp();
    const float half = fmul(sm.acc[l][0][num - 1], 0.5f);  // wSum over the ascending-depth order
    // weighted medians: first sorted position whose running sum reaches half (0 depth, 1..3 normal components)
    int hit[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int first = 1 << 20;
#pragma unroll
        for (int r = 3; r >= 0; r--) {
            const int s_ = lane * 4 + r;
            if (s_ < num && sm.acc[l][q][s_] >= half) first = s_;
        }
        first = __reduce_min_sync(0xffffffffu, first);
        hit[q] = first < num ? (int)sm.slot[l][q][first] : -1;
    }
    if (lane == 0) {
        const int hd = hit[0], hx = hit[1], hy = hit[2], hz = hit[3];
        float4 norm_mid = make_float4(0.f, 0.f, 0.f, 0.f);  // (uninitialised in the reference when never reached)
        if (hx >= 0) norm_mid.x = sm.x[l][hx];
        if (hy >= 0) norm_mid.y = sm.y[l][hy];
        if (hz >= 0) norm_mid.z = sm.z[l][hz];
        if (hd >= 0) {
            // weimid.  The dummy element can be the one that reaches half: with a single reliable neighbour of positive
            // depth the sorted list is (dummy, neighbour), only position 0 counts, wSum = 0.  Its index is 0 in the
            // reference's zero-initialised list, so the reference then takes pixel (0, 0) and ITS depth.
            const bool dummy = hd == kCoopDummy;
            const int ii = hd / 11, jj = hd - ii * 11;
            const int mx = dummy ? 0 : px - radius + ii * gap, my = dummy ? 0 : y - radius + jj * gap;
            const float disp_mid = fdiv(fmul(g.f_params, g.baseline), dummy ? depth[0] : sm.d[l][hd]);
            const float len = __fsqrt_rn(dot3(norm_mid.x, norm_mid.x, norm_mid.y, norm_mid.y, norm_mid.z, norm_mid.z));
            norm_mid.x = fdiv(norm_mid.x, len); norm_mid.y = fdiv(norm_mid.y, len); norm_mid.z = fdiv(norm_mid.z, len);
            norm_mid.w = g_plane_d(g, norm_mid.x, norm_mid.y, norm_mid.z, mx, my, disp_mid);
        }
        const int ths = 24 / po;
        const float fb = fmul(g.f_params, g.baseline);
        const float disp_now = fdiv(fb, g_plane_depth(g, norm_mid, px, y));
        const float disp_org = fdiv(fb, g_plane_depth(g, plane[p], px, y));
        scale_out[p] = (fabsf(fsub(disp_now, disp_org)) > (float)ths) ? 0.f : 1.f;
    }
}

// snapshot = pre-launch copies of the arrays the launch rewrites (scratch supplied by the context)
static inline int wmf_launch(const GlueConst &g, const float *ref, float4 *plane, float *depth, float *scale, float *scale_snapshot,
                             int iter, bool per_thread, cudaStream_t s) {
    const size_t n = (size_t)g.W * g.H;
    if (cudaMemcpyAsync(scale_snapshot, scale, n * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return TSAR_ERR_CUDA;
    if (per_thread) {  // the first implementation, kept for A/B checks (TSAR_B200_WMF_PER_THREAD=1)
        dim3 b(32, 2), grid((g.W + 31) / 32, (g.H + 1) / 2);
        wmf_kernel<<<grid, b, 0, s>>>(g, ref, plane, depth, scale_snapshot, scale, iter);
    } else {
        if (cudaFuncSetAttribute(wmf_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WmfCoopSmem)) != cudaSuccess)
            return TSAR_ERR_CUDA;
        dim3 grid((g.W + kCoopRun - 1) / kCoopRun, g.H);
        wmf_coop_kernel<<<grid, 256, sizeof(WmfCoopSmem), s>>>(g, ref, plane, depth, scale_snapshot, scale, iter);
    }
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
