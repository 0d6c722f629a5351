// glue_kernels.cuh -- confidence, TSAR glue and depth-completion kernels (north-star item 4 and the
// confidence part of item 3).  All are one-thread-per-pixel maps; they are HBM-bound (bytes per pixel
// in DESIGN.md).  Reference: gipuma.cu:732-755, 810-844, 1161-1292.
#pragma once
#include "pm_core.cuh"

namespace tsar {

struct CamDev {  // per image (not per selected view): needed by the reverse cost, which indexes by beview
    float R[9];
    float t[3];
    float K[9];
};

struct GlueConst {
    int W, H;
    int hrad, vrad;
    float Kinv[9], Minv[9], Pc[3], C[3];
    float Rorig[9], Rorig_inv[9];
    float fx, alpha, cx, cy, f_params, baseline;
    float min_disp, max_disp, depthMin, depthMax;
};

// depth of plane along the pixel ray / plane offset / view vector, on GlueConst
__device__ __forceinline__ float g_plane_depth(const GlueConst &g, const float4 &pl, int x, int y) {
    if (pl.w != pl.w) return 1000.0f;
    const float dy = fsub((float)y, g.cy), dx = fsub((float)x, g.cx);
    const float den = ffma(pl.z, g.fx, ffma(pl.x, dx, fmul(g.alpha, fmul(pl.y, dy))));
    return fdiv(fmul(pl.w, -g.fx), den);
}
__device__ __forceinline__ float g_plane_d(const GlueConst &g, float nx, float ny, float nz, int x, int y, float depth) {
    const float ptx = ffma((float)x, depth, -g.Pc[0]);
    const float pty = ffma((float)y, depth, -g.Pc[1]);
    const float ptz = fsub(depth, g.Pc[2]);
    float X, Y, Z;
    matvec3(g.Minv, ptx, pty, ptz, X, Y, Z);
    return -dot3(nx, X, ny, Y, nz, Z);
}
__device__ __forceinline__ void g_view_vector(const GlueConst &g, int x, int y, float &vx, float &vy, float &vz) {
    const float ptx = fsub((float)x, g.Pc[0]), pty = fsub((float)y, g.Pc[1]), ptz = fsub(1.0f, g.Pc[2]);
    float X, Y, Z;
    matvec3(g.Minv, ptx, pty, ptz, X, Y, Z);
    vx = fsub(X, g.C[0]); vy = fsub(Y, g.C[1]); vz = fsub(Z, g.C[2]);
    const float r = rsqrtf(dot3(vx, vx, vy, vy, vz, vz));
    vx = fmul(vx, r); vy = fmul(vy, r); vz = fmul(vz, r);
}

// ---------------------------------------------------------------------------------------------
// gipuma_getlrdiff + rlCost (gipuma.cu:1161-1186, 301-392): reverse (source -> reference) bilateral
// NCC in the best view, through the inverse homography (adjugate / determinant).
// ---------------------------------------------------------------------------------------------
__global__ void lrdiff_kernel(const __grid_constant__ GlueConst g, const CamDev *__restrict__ cams,
                              const cudaTextureObject_t *__restrict__ tex, int n_images,
                              const float4 *__restrict__ plane, const float *__restrict__ cost,
                              const int *__restrict__ beview, float *__restrict__ lrdiff) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const size_t p = (size_t)y * g.W + x;
    int to = beview[p];
    if (to < 0 || to >= n_images) to = 0;  // the reference would read gs.imgs[-1]; beview is never -1 after
                                            // an accepted update, keep memory-safe
    const CamDev cam = cams[to];
    const float4 pl = plane[p];
    // forward homography (same arithmetic as the PatchMatch path)
    float Hm[9];
    {
        float A[9], T[9];
        const float n[3] = {pl.x, pl.y, pl.z};
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int q = 0; q < 3; q++) A[r * 3 + q] = fsub(cam.R[r * 3 + q], fdiv(fmul(cam.t[r], n[q]), pl.w));
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int q = 0; q < 3; q++)
                T[r * 3 + q] = dot3(A[r * 3], g.Kinv[q], A[r * 3 + 1], g.Kinv[3 + q], A[r * 3 + 2], g.Kinv[6 + q]);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int q = 0; q < 3; q++)
                Hm[r * 3 + q] = dot3(cam.K[r * 3], T[q], cam.K[r * 3 + 1], T[3 + q], cam.K[r * 3 + 2], T[6 + q]);
    }
    // determinant and adjugate as the reference build evaluates them (products of two are rounded and
    // shared between det and the 2x2 minors; third factors / second products are fused)
    const float h04 = fmul(Hm[0], Hm[4]), h15 = fmul(Hm[1], Hm[5]), h23 = fmul(Hm[2], Hm[3]);
    const float h24 = fmul(Hm[2], Hm[4]), h13 = fmul(Hm[1], Hm[3]), h05 = fmul(Hm[0], Hm[5]);
    float det = fmul(Hm[8], h04);
    det = ffma(Hm[6], h15, det);
    det = ffma(Hm[7], h23, det);
    det = ffma(Hm[6], -h24, det);
    det = ffma(Hm[8], -h13, det);
    det = ffma(Hm[7], -h05, det);
    float V[9];
    V[0] = ffma(Hm[4], Hm[8], -fmul(Hm[5], Hm[7]));
    V[1] = ffma(Hm[1], Hm[8], -fmul(Hm[2], Hm[7]));
    V[2] = fsub(h15, h24);
    V[3] = ffma(Hm[3], Hm[8], -fmul(Hm[5], Hm[6]));
    V[4] = ffma(Hm[0], Hm[8], -fmul(Hm[2], Hm[6]));
    V[5] = fsub(h05, h23);
    V[6] = ffma(Hm[3], Hm[7], -fmul(Hm[4], Hm[6]));
    V[7] = ffma(Hm[0], Hm[7], -fmul(Hm[1], Hm[6]));
    V[8] = fsub(h04, h13);
    V[0] = fdiv(V[0], det);  V[1] = fdiv(-V[1], det); V[2] = fdiv(V[2], det);
    V[3] = fdiv(-V[3], det); V[4] = fdiv(V[4], det);  V[5] = fdiv(-V[5], det);
    V[6] = fdiv(V[6], det);  V[7] = fdiv(-V[7], det); V[8] = fdiv(V[8], det);

    // centre of the window in the source view
    const float fx0 = (float)x, fy0 = (float)y;
    const float cz = fadd(ffma(Hm[6], fx0, fmul(Hm[7], fy0)), Hm[8]);
    const float pcx = fdiv(fadd(ffma(Hm[0], fx0, fmul(Hm[1], fy0)), Hm[2]), cz);
    const float pcy = fdiv(fadd(ffma(Hm[3], fx0, fmul(Hm[4], fy0)), Hm[5]), cz);
    const cudaTextureObject_t tl = tex[0], tr = tex[to];
    const float cen = tex2D<float>(tr, fadd(pcx, 0.5f), fadd(pcy, 0.5f));

    float sum_ref = 0.f, sum_rr = 0.f, sum_src = 0.f, sum_ss = 0.f, sum_rs = 0.f, wsum = 0.f;
    for (int i = -g.hrad; i < g.hrad + 1; i += 2) {
        const int plx = (int)fadd(pcx, (float)i);  // make_int2(pt_c.x + i, ..): float add, truncation
        const float fplx = (float)plx;
        for (int j = -g.vrad; j < g.vrad + 1; j += 2) {
            const int ply = (int)fadd(pcy, (float)j);
            const float fply = (float)ply;
            const float ref_pix = tex2D<float>(tr, fadd(fplx, 0.5f), fadd(fply, 0.5f));
            const float Z = fadd(ffma(V[6], fplx, fmul(V[7], fply)), V[8]);
            const float X = fdiv(fadd(ffma(V[0], fplx, fmul(V[1], fply)), V[2]), Z);
            const float Y = fdiv(fadd(ffma(V[3], fplx, fmul(V[4], fply)), V[5]), Z);
            const float src_pix = tex2D<float>(tl, fadd(X, 0.5f), fadd(Y, 0.5f));
            const float w = expf(fsub(spatial_term(i, j), fdiv(fabsf(fsub(ref_pix, cen)), 18.0f)));
            const float wr = fmul(ref_pix, w), ws = fmul(src_pix, w);
            sum_ref = fadd(sum_ref, wr);
            sum_rr = ffma(ref_pix, wr, sum_rr);
            sum_src = fadd(sum_src, ws);
            sum_ss = ffma(src_pix, ws, sum_ss);
            sum_rs = ffma(src_pix, wr, sum_rs);
            wsum = fadd(wsum, w);
        }
    }
    const float inv = __frcp_rn(wsum);
    const float sr = fmul(inv, sum_ref), ss = fmul(inv, sum_src);
    const float var_ref = ffma(inv, sum_rr, -fmul(sr, sr));
    const float var_src = ffma(inv, sum_ss, -fmul(ss, ss));
    float rcost;
    if (fminf(var_ref, var_src) < 1e-5f) rcost = kMaxCost;
    else {
        const float covar = ffma(-sr, ss, fmul(inv, sum_rs));
        rcost = fmaxf(0.0f, fminf(kMaxCost, fsub(1.0f, fdiv(covar, __fsqrt_rn(fmul(var_ref, var_src))))));
    }
    float d = fabsf(fsub(cost[p], rcost));  // gipuma.cu:1182-1185
    if (d > 1.0f) d = 1.0f;
    lrdiff[p] = d;
}

// gipuma_getview (gipuma.cu:1189-1213): confidence + disparity of the current plane.  12 B read, 8 B written/px (+16 B plane)
__global__ void getview_kernel(const __grid_constant__ GlueConst g, const float4 *__restrict__ plane,
                               const float *__restrict__ cost, const float *__restrict__ lrdiff,
                               float *__restrict__ confid, float *__restrict__ depth) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const size_t p = (size_t)y * g.W + x;
    // ((2-c)/2 + (1-lrdiff))/2, as compiled: fma(2-c, 0.5, 1-lrdiff) * 0.5
    confid[p] = fmul(ffma(fsub(2.0f, cost[p]), 0.5f, fsub(1.0f, lrdiff[p])), 0.5f);
    const float dep = g_plane_depth(g, plane[p], x, y);
    depth[p] = fdiv(fmul(g.f_params, g.baseline), dep);
}

// gipuma_get_disp (gipuma.cu:732-755): imported world normals -> reference camera frame, plane offset
// from the imported disparity.
__global__ void get_disp_kernel(const __grid_constant__ GlueConst g, float4 *__restrict__ plane,
                                const float *__restrict__ depth) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const size_t p = (size_t)y * g.W + x;
    const float4 n = plane[p];
    float nx, ny, nz;
    matvec3(g.Rorig, n.x, n.y, n.z, nx, ny, nz);
    const float dep = fdiv(fmul(g.f_params, g.baseline), depth[p]);
    plane[p] = make_float4(nx, ny, nz, g_plane_d(g, nx, ny, nz, x, y, dep));
}

// gipuma_compute_disp (gipuma.cu:810-844): output layout -- world normal in xyz, depth in w (0 if c == MAXCOST)
__global__ void compute_disp_kernel(const __grid_constant__ GlueConst g, float4 *__restrict__ plane,
                                    const float *__restrict__ cost) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const size_t p = (size_t)y * g.W + x;
    const float4 n = plane[p];
    float4 o;
    matvec3(g.Rorig_inv, n.x, n.y, n.z, o.x, o.y, o.z);
    o.w = (cost[p] != kMaxCost) ? g_plane_depth(g, n, x, y) : 0.0f;
    plane[p] = o;
}

// gipuma_update_scale (gipuma.cu:1216-1259): depth completion -- every pixel of a textureless region
// takes the region plane (flipped to face the camera); all pixels get depth[] refreshed.
__global__ void update_scale_kernel(const __grid_constant__ GlueConst g, float4 *__restrict__ plane,
                                    float *__restrict__ cost, float *__restrict__ scale,
                                    float *__restrict__ depth, const float *__restrict__ canny,
                                    const float *__restrict__ region_text, const float4 *__restrict__ region_plane,
                                    int n_regions) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const size_t p = (size_t)y * g.W + x;
    const int reg = (int)canny[p];  // label stored as float
    float4 pl = plane[p];
    if (reg >= 0 && reg < n_regions && region_text[reg] == -1.0f) {
        cost[p] = 0.0f;
        scale[p] = 1.0f;
        float vx, vy, vz;
        g_view_vector(g, x, y, vx, vy, vz);
        pl = region_plane[reg];
        if (dot3(pl.x, vx, pl.y, vy, pl.z, vz) > 0.0f) { pl.x = -pl.x; pl.y = -pl.y; pl.z = -pl.z; pl.w = -pl.w; }
        plane[p] = pl;
    }
    const float dep = g_plane_depth(g, pl, x, y);
    depth[p] = fdiv(fmul(g.f_params, g.baseline), dep);
}

// gipuma_update_scale_2 (gipuma.cu:1262-1292): fakedepth <- depth of the region plane, textureless regions only
__global__ void update_scale_2_kernel(const __grid_constant__ GlueConst g, float *__restrict__ fakedepth,
                                      const float *__restrict__ canny, const float *__restrict__ region_text,
                                      const float4 *__restrict__ region_plane, int n_regions) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const size_t p = (size_t)y * g.W + x;
    const int reg = (int)canny[p];
    if (reg >= 0 && reg < n_regions && region_text[reg] == -1.0f) {
        float vx, vy, vz;
        g_view_vector(g, x, y, vx, vy, vz);
        float4 pl = region_plane[reg];
        if (dot3(pl.x, vx, pl.y, vy, pl.z, vz) > 0.0f) { pl.x = -pl.x; pl.y = -pl.y; pl.z = -pl.z; pl.w = -pl.w; }
        fakedepth[p] = g_plane_depth(g, pl, x, y);
    }
}

// canny[] from the quarter-resolution region labels (main.cpp:558-568): label of (y/4, x/4), the index stepped
// back once where the twice-halved size is smaller than ceil(size/4).  Labels stay float at the boundary (Q13).
__global__ void labels_quarter_kernel(const int *__restrict__ lab, int wq, int hq, int W, int H, float *__restrict__ canny) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    int sx = x / 4, sy = y / 4;
    if (sx >= wq) sx--;
    if (sy >= hq) sy--;
    canny[(size_t)y * W + x] = (float)lab[(size_t)sy * wq + sx];
}

// lines->scale from the confidence map: 1 where confid > thr.  The shipped flow takes the reliable flags from APD's weak.png
// (main.cpp:1499-1514); when PatchMatch runs inside the library there is no such file and the confidence of
// gipuma_getview stands in for it (tsar_scale_from_confidence).
__global__ void scale_from_confidence_kernel(const float *__restrict__ confid, float thr, float *__restrict__ scale, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) scale[i] = confid[i] > thr ? 1.0f : 0.0f;
}

// lines->scale from APD's weak.png (main.cpp:1499-1514): 1 where the BGR pixel is white, green (0,255,0) or red (0,0,255);
// other pixels keep their value (the array is zero after LineState::resize)
__global__ void scale_from_weak_png_kernel(const uchar3 *__restrict__ bgr, float *__restrict__ scale, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uchar3 p = bgr[i];
    if ((p.x == 255 && p.y == 255 && p.z == 255) || (p.x == 0 && p.y == 255 && p.z == 0) || (p.x == 0 && p.y == 0 && p.z == 255)) scale[i] = 1.0f;
}

// Output files hold the depth and the normals in separate arrays (TSAR_disp.dmb: w of the output layout, TSAR_normals.dmb:
// xyz, main.cpp:1785-1795): split on the device so that the host receives exactly the file payloads.
__global__ void split_outputs_kernel(const float4 *__restrict__ out4, float *__restrict__ depth, float *__restrict__ normals, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = out4[i];
    depth[i] = v.w;
    normals[3 * i] = v.x; normals[3 * i + 1] = v.y; normals[3 * i + 2] = v.z;
}

}  // namespace tsar
