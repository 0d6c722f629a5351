// pm_core.cuh -- device-side building blocks of the PatchMatch depthmap path (sm_100a).
//
// What the reference computes (gipuma.cu, file:line in each function) is kept, how it is computed is
// not:
//   * every term of the bilateral-NCC window that depends only on the reference image -- the
//     weights w_k, the products w_k*ref_k, sum(w), sum(w*ref), sum(w*ref^2) -- is computed ONCE per
//     pixel per launch and staged in shared memory ([sample][thread] float2, conflict-free LDS.64),
//     instead of once per (plane hypothesis x source view) as pmCost does (gipuma.cu:259-277).  The
//     per-accumulator summation order is unchanged, so the sums are bit-identical;
//   * camera constants live in the kernel-parameter constant bank, not behind 8 managed pointers
//     per camera;
//   * the 9 + 2*S IEEE divisions per cost evaluation share one refined reciprocal per divisor;
//   * random numbers come from a per-launch table of the XORWOW output stream of each image row
//     (curand_init(seed, y, x) == stream of row y advanced x draws), not from a per-pixel
//     curand_init skip-ahead;
//   * same-colour neighbours are read from the pre-launch snapshot (deterministic), see DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pm_math.cuh"

// outer-loop unroll of the branch-free sampling loop (window columns per trip); a translation unit may
// override it before including this header
#ifndef PM_FAST_UNROLL
#define PM_FAST_UNROLL(n1) (((n1) + 1) / 2)
#endif

namespace tsar {

constexpr int kMaxViews = 32;
constexpr float kMaxCost = 2.0f;  // config.h:22 MAXCOST

struct ViewC {      // one source view, in the frame of the reference camera
    float R[9];     // Camera_cu::R
    float t[3];     // Camera_cu::t4
    float K[9];     // Camera_cu::K (own intrinsics)
};

struct PmConst {
    int W, H, V;
    int hrad, vrad;        // (box-1)/2, gipuma.cu:858-859
    int n1x, n1y, ns;      // samples per axis (stride 2) and per window
    int n_best, cost_comb;
    int y_limit;           // rows the reference's checkerboard grid reaches (gipuma.cu:1721)
    int rng_pitch;
    int k_pinhole;         // 1: all intrinsics (and K_ref^-1) are zero-skew pinhole matrices with a unit last row (host-side
                           // dispatch to the PIN = true instantiation of the checkerboard kernel)
    float Kinv[9];         // cameras[0].K_inv
    float Minv[9];         // cameras[0].M_inv
    float Pc[3];           // cameras[0].P_col34
    float C[3];            // cameras[0].C4
    float fx, alpha, cx, cy;   // cameras[0].fx, .alpha, .K[2], .K[5]
    float f_params;        // CameraParameters_cu::f
    float f_cam0;          // cameras[0].f
    float baseline;        // cameras[0].baseline
    float depthMin, depthMax;
    float min_disp, max_disp;
    cudaTextureObject_t tex[kMaxViews];  // source view textures (fp32 texels), order of viewSelectionSubset
    cudaTextureObject_t tex8[kMaxViews]; // the same views as 8-bit unorm textures (valid when every image is 8-bit valued)
    int view_id[kMaxViews];              // viewSelectionSubset[i]
    ViewC view[kMaxViews];
};

// ---------------------------------------------------------------------------------------------
// window weights (reference-only part of pmCost, gipuma.cu:247, 259-277)
// ---------------------------------------------------------------------------------------------
struct RefStats {
    float inv;      // 1 / sum(w)
    float sr;       // sum(w*ref) * inv
    float var_ref;  // sum(w*ref^2)*inv - sr^2
};

// spatial term of the bilateral weight: -sqrt(i^2+j^2) / (2*5*5), as compiled: sqrt.rn / -50
__device__ __forceinline__ float spatial_term(int i, int j) {
    return fdiv(__fsqrt_rn((float)(i * i + j * j)), -50.0f);
}

// Where the hoisted per-sample terms of a thread live in shared memory.
//  WPair: (w, w*ref) as one float2 per sample per thread -- [ns][NT] table, one LDS.64 per sample.
//  WTile: w alone per sample per thread ([ns][NT] floats) + ONE reference-image tile per CTA; w*ref is re-formed by
//         the same single multiplication.  Half the bytes per thread: what lets the 100-sample (19x19) window run
//         3 CTAs per SM instead of 2.  Pitch 64 floats: lanes of a warp differ in column (and, on the checkerboard,
//         by one row = 64 floats), so every access is bank-conflict free.
template <int NT>
struct WPair {
    float2 *p;
    __device__ __forceinline__ void put(int k, float w, float tr) const { p[k * NT] = make_float2(w, tr); }
    __device__ __forceinline__ float2 get(int k, int, int) const { return p[k * NT]; }
};
constexpr int kTilePitch = 64;
template <int NT>
struct WTile {
    float *w;
    const float *rt;  // tile element of the thread's window origin (x - hrad, y - vrad)
    __device__ __forceinline__ void put(int k, float wv, float) const { w[k * NT] = wv; }
    __device__ __forceinline__ float2 get(int k, int ii, int jj) const {
        const float wv = w[k * NT];
        return make_float2(wv, fmul(rt[2 * jj * kTilePitch + 2 * ii], wv));
    }
};

// N1 = samples per axis known at compile time (hRad+1, square window) or 0 for run-time sizes.
template <int NT, int N1, class WS>
__device__ __forceinline__ RefStats window_weights(const PmConst &c, const float *__restrict__ ref, int x, int y,
                                                   const float *__restrict__ sp, const WS &ws) {
    const int W = c.W, H = c.H;
    // tex2D(l, x+0.5, y+0.5) with unnormalised coords: clamp addressing, exact texel
    const int xc = min(max(x, 0), W - 1), yc = min(max(y, 0), H - 1);
    const float cen = __ldg(ref + (size_t)yc * W + xc);
    float sum_ref = 0.f, sum_rr = 0.f, wsum = 0.f;
    const int n1x = N1 ? N1 : c.n1x, n1y = N1 ? N1 : c.n1y;
    const int hrad = N1 ? N1 - 1 : c.hrad, vrad = N1 ? N1 - 1 : c.vrad;
    int k = 0;
    // runs once per pixel per launch: kept rolled (code size matters more than its speed)
#pragma unroll 1
    for (int ii = 0; ii < n1x; ii++) {
        const int xi = min(max(x - hrad + 2 * ii, 0), W - 1);
#pragma unroll 2
        for (int jj = 0; jj < n1y; jj++, k++) {
            const int yj = min(max(y - vrad + 2 * jj, 0), H - 1);
            const float r = __ldg(ref + (size_t)yj * W + xi);
            // exp(-spatial/50 - |ref-cen|/18), gipuma.cu:266-268
            const float w = expf(fsub(sp[k], fdiv(fabsf(fsub(r, cen)), 18.0f)));
            const float tr = fmul(r, w);
            sum_ref = fadd(sum_ref, tr);       // gipuma.cu:270
            sum_rr = ffma(r, tr, sum_rr);      // gipuma.cu:271
            wsum = fadd(wsum, w);              // gipuma.cu:275
            ws.put(k, w, tr);
        }
    }
    RefStats s;
    s.inv = __frcp_rn(wsum);                               // gipuma.cu:279
    s.sr = fmul(s.inv, sum_ref);                           // :280
    s.var_ref = ffma(s.inv, sum_rr, -fmul(s.sr, s.sr));    // :281,286
    return s;
}

// ---------------------------------------------------------------------------------------------
// plane-induced homography H = K_src (R - t n^T / d) K_ref^-1   (getHomography_cu, gipuma.cu:207-224)
// ---------------------------------------------------------------------------------------------
template <bool PIN>
__device__ __forceinline__ void homography(const PmConst &c, const ViewC &v, const float4 &pl, float *__restrict__ Hm) {
    float A[9];
    const float n[3] = {pl.x, pl.y, pl.z};
    // outer product t n^T, each entry / d (9 IEEE divisions by the same d), R - .
    // |t_r * n_q| is far below 2^60 for any camera; tiny numerators only perturb R below its ulp
    if (div_range_ok(fabsf(pl.w))) {
        const float r = refined_rcp(pl.w);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int q = 0; q < 3; q++) A[i * 3 + q] = fsub(v.R[i * 3 + q], div_refined(fmul(v.t[i], n[q]), pl.w, r));
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int q = 0; q < 3; q++) A[i * 3 + q] = fsub(v.R[i * 3 + q], fdiv(fmul(v.t[i], n[q]), pl.w));
    }
    float T[9];
    if (PIN) {
        // Every intrinsic matrix is [[fx,0,cx],[0,fy,cy],[0,0,1]] (and K_ref^-1 has the same zero pattern): the products
        // with the exact zeros contribute +-0 and the one with 1 is exact, so they are dropped.  The remaining
        // operations are the reference's, in its order; only the sign of an exactly-zero entry can differ, which no
        // later operation observes (entries are only multiplied by pixel coordinates and added to non-zero terms, or
        // compared by magnitude).
#pragma unroll
        for (int r = 0; r < 3; r++) {
            T[r * 3 + 0] = fmul(A[r * 3 + 0], c.Kinv[0]);
            T[r * 3 + 1] = fmul(A[r * 3 + 1], c.Kinv[4]);
            T[r * 3 + 2] = dot3(A[r * 3 + 0], c.Kinv[2], A[r * 3 + 1], c.Kinv[5], A[r * 3 + 2], c.Kinv[8]);
        }
#pragma unroll
        for (int q = 0; q < 3; q++) {
            Hm[0 * 3 + q] = ffma(v.K[2], T[2 * 3 + q], fmul(v.K[0], T[0 * 3 + q]));
            Hm[1 * 3 + q] = ffma(v.K[5], T[2 * 3 + q], fmul(v.K[4], T[1 * 3 + q]));
            Hm[2 * 3 + q] = T[2 * 3 + q];
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++)
            T[r * 3 + q] = dot3(A[r * 3 + 0], c.Kinv[0 * 3 + q], A[r * 3 + 1], c.Kinv[1 * 3 + q], A[r * 3 + 2],
                                c.Kinv[2 * 3 + q]);
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++)
            Hm[r * 3 + q] = dot3(v.K[r * 3 + 0], T[0 * 3 + q], v.K[r * 3 + 1], T[1 * 3 + q], v.K[r * 3 + 2],
                                 T[2 * 3 + q]);
}

// ---------------------------------------------------------------------------------------------
// one pmCost (gipuma.cu:230-298) with the reference-only terms hoisted
// ---------------------------------------------------------------------------------------------
// PXF selects how H * (px, py, 1) is rounded (config.h matvecmul4noz, m0*px + m1*py + m2):
//   false: fadd(m2, fma(m0, px, m1*py))  -- what nvcc emits inside the reference's real kernels
//          (gipuma_init_cu2, *_spatialProp_cu, *_planeRefine_cu; read from their SASS);
//   true:  fadd(m2, fma(m1, py, m0*px))  -- the loop-hoisted form nvcc happens to emit for the oracle's
//          stand-alone wrapper kernel (oracle/ref_driver.cu: ref_eval_kernel).  Test-only switch so the
//          unit-level parity test can demand bit equality against that wrapper.
//
// U8: fetch from the 8-bit copy of the source view.  The texture unit's bilinear result for 8-bit texels,
// read as normalised float v, satisfies rint(v * 255 * 256) == N with N / 256 exactly the fp32-texel result
// (weights have 8 fractional bits; calibrated on B200: 400k samples incl. borders, 100 % identical).  The
// window sums are carried in units of N (i.e. scaled by 2^8 / 2^16, exact in binary floating point) and
// scaled back once per evaluation, so costs are bit-identical to the fp32-texture path while each warp-wide
// fetch moves a quarter of the texel bytes through the L1TEX data pipe (the measured limiter).
template <int NT, int N1, bool PXF, bool U8, bool PIN, class WS>
__device__ __forceinline__ float view_cost(const PmConst &c, int vi, int x, int y, const float4 &pl,
                                           const WS &ws, const RefStats &rs) {
    float Hm[9];
    homography<PIN>(c, c.view[vi], pl, Hm);
    const cudaTextureObject_t tex = U8 ? c.tex8[vi] : c.tex[vi];
    float s_s = 0.f, s_ss = 0.f, s_rs = 0.f;
    const int n1x = N1 ? N1 : c.n1x, n1y = N1 ? N1 : c.n1y;
    const int hrad = N1 ? N1 - 1 : c.hrad, vrad = N1 ? N1 - 1 : c.vrad;

    // Can every x/z, y/z of this window take the branch-free division?  X, Y, Z are affine in the sample
    // position, so they are bounded by their values at the four window corners; demand that z keeps its
    // sign, does not cancel (|z| >= 2^-10 of its term magnitudes) and that everything is inside
    // 2^-60..2^60.  NaN anywhere fails the test and takes the exact path.
    const float xl = (float)(x - hrad), xr = (float)(x - hrad + 2 * (n1x - 1));
    const float yt = (float)(y - vrad), yb = (float)(y - vrad + 2 * (n1y - 1));
    const float axm = fmaxf(fabsf(xl), fabsf(xr)), aym = fmaxf(fabsf(yt), fabsf(yb));
    const float zs = fabsf(Hm[6]) * axm + fabsf(Hm[7]) * aym + fabsf(Hm[8]);
    const float xs = fabsf(Hm[0]) * axm + fabsf(Hm[1]) * aym + fabsf(Hm[2]);
    const float ys = fabsf(Hm[3]) * axm + fabsf(Hm[4]) * aym + fabsf(Hm[5]);
    const float z00 = Hm[6] * xl + Hm[7] * yt + Hm[8], z10 = Hm[6] * xr + Hm[7] * yt + Hm[8];
    const float z01 = Hm[6] * xl + Hm[7] * yb + Hm[8], z11 = Hm[6] * xr + Hm[7] * yb + Hm[8];
    const float zlo = fminf(fminf(z00, z10), fminf(z01, z11)), zhi = fmaxf(fmaxf(z00, z10), fmaxf(z01, z11));
    const float zmin = (zlo > 0.0f) ? zlo : ((zhi < 0.0f) ? -zhi : 0.0f);
    const bool fast = (zmin >= kDivLo) && (zmin >= 9.765625e-4f * zs) && (zs <= kDivHi) && (xs <= kDivHi) && (ys <= kDivHi);

    int k = 0;
    constexpr int kUnr = N1 ? PM_FAST_UNROLL(N1) : 1;
    if (fast) {
        // half of the window per trip when the size is known: enough independent texture fetches in flight
        // to cover their latency, small enough to stay in the instruction cache
#pragma unroll(kUnr)
        for (int ii = 0; ii < n1x; ii++) {
            const float px = (float)(x - hrad + 2 * ii);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            if (PXF) { a0 = fmul(Hm[0], px); a1 = fmul(Hm[3], px); a2 = fmul(Hm[6], px); }
#pragma unroll
            for (int jj = 0; jj < n1y; jj++, k++) {
                const float py = (float)(y - vrad + 2 * jj);
                const float X = fadd(Hm[2], PXF ? ffma(Hm[1], py, a0) : ffma(Hm[0], px, fmul(Hm[1], py)));
                const float Y = fadd(Hm[5], PXF ? ffma(Hm[4], py, a1) : ffma(Hm[3], px, fmul(Hm[4], py)));
                const float Z = fadd(Hm[8], PXF ? ffma(Hm[7], py, a2) : ffma(Hm[6], px, fmul(Hm[7], py)));
                const float r = refined_rcp(Z);
                const float xs_ = fadd(div_refined(X, Z, r), 0.5f);
                const float ys_ = fadd(div_refined(Y, Z, r), 0.5f);
                float src = tex2D<float>(tex, xs_, ys_);  // hardware bilinear, clamp
                if (U8) src = fsub(ffma(src, 65280.0f, 12582912.0f), 12582912.0f);  // N = rint(v * 255 * 256)
                const float2 w = ws.get(k, ii, jj);            // (w, w*ref)
                const float ts = fmul(src, w.x);
                s_s = fadd(s_s, ts);            // gipuma.cu:272
                s_ss = ffma(src, ts, s_ss);     // :273
                s_rs = ffma(src, w.y, s_rs);    // :274
            }
        }
    } else {
        // exact IEEE divisions (rare: plane nearly through the camera centre, window crossing z = 0, ...)
        for (int ii = 0; ii < n1x; ii++) {
            const float px = (float)(x - hrad + 2 * ii);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            if (PXF) { a0 = fmul(Hm[0], px); a1 = fmul(Hm[3], px); a2 = fmul(Hm[6], px); }
            for (int jj = 0; jj < n1y; jj++, k++) {
                const float py = (float)(y - vrad + 2 * jj);
                const float X = fadd(Hm[2], PXF ? ffma(Hm[1], py, a0) : ffma(Hm[0], px, fmul(Hm[1], py)));
                const float Y = fadd(Hm[5], PXF ? ffma(Hm[4], py, a1) : ffma(Hm[3], px, fmul(Hm[4], py)));
                const float Z = fadd(Hm[8], PXF ? ffma(Hm[7], py, a2) : ffma(Hm[6], px, fmul(Hm[7], py)));
                float src = tex2D<float>(tex, fadd(fdiv(X, Z), 0.5f), fadd(fdiv(Y, Z), 0.5f));
                if (U8) src = fsub(ffma(src, 65280.0f, 12582912.0f), 12582912.0f);
                const float2 w = ws.get(k, ii, jj);
                const float ts = fmul(src, w.x);
                s_s = fadd(s_s, ts);
                s_ss = ffma(src, ts, s_ss);
                s_rs = ffma(src, w.y, s_rs);
            }
        }
    }
    if (U8) { s_s = fmul(s_s, 0.00390625f); s_rs = fmul(s_rs, 0.00390625f); s_ss = fmul(s_ss, 1.52587890625e-05f); }
    const float ss = fmul(rs.inv, s_s);
    const float var_src = ffma(rs.inv, s_ss, -fmul(ss, ss));
    const float rsn = fmul(rs.inv, s_rs);
    if (fminf(rs.var_ref, var_src) < 1e-5f) return kMaxCost;  // gipuma.cu:289-291
    const float covar = ffma(-rs.sr, ss, rsn);
    const float denom = __fsqrt_rn(fmul(rs.var_ref, var_src));
    return fmaxf(0.0f, fminf(kMaxCost, fsub(1.0f, fdiv(covar, denom))));  // :295
}

// ---------------------------------------------------------------------------------------------
// pmCostMultiview_cu (gipuma.cu:456-518)
// ---------------------------------------------------------------------------------------------
struct MvResult {
    float cost;
    float ratio;
    int beview;
};

// GENERIC = false: only the two smallest costs are ever read (cost_comb == COMB_BEST_N and
// n_best <= 2, the setting of every run script), kept in registers.  GENERIC = true: any n_best /
// COMB_ALL through the reference's full insertion sort (local-memory arrays, as the reference).
template <int NT, int N1, bool GENERIC, bool PXF = false, bool U8 = false, bool PIN = false, class WS = WPair<NT>>
__device__ __forceinline__ MvResult multiview_cost(const PmConst &c, int x, int y, const float4 &pl,
                                                   const WS &ws, const RefStats &rs) {
    MvResult out;
    if (!GENERIC) {
        float s0 = __int_as_float(0x7f800000), s1 = __int_as_float(0x7f800000);
        int nvalid = 0, bidx = -1;
        for (int vi = 0; vi < c.V; vi++) {
            float cv = view_cost<NT, N1, PXF, U8, PIN>(c, vi, x, y, pl, ws, rs);
            if (cv < kMaxCost) nvalid++;
            else cv = kMaxCost;
            if (cv < s0) { s1 = s0; s0 = cv; bidx = vi; }
            else {
                if (cv == s0) bidx = vi;      // "last view whose cost equals the minimum", gipuma.cu:506-510
                if (cv < s1) s1 = cv;
            }
        }
        const int nb = min(nvalid, c.n_best);
        if (nb > 0) {
            out.cost = (nb == 1) ? s0 : fdiv(fadd(s0, s1), 2.0f);
            out.ratio = fdiv(s0, s1);
            out.beview = c.view_id[bidx];
        } else {
            out.cost = kMaxCost; out.ratio = 0.f; out.beview = -1;
        }
        return out;
    } else {
        // general combination: full insertion sort as the reference does (sort_small, gipuma.cu:425-434)
        float cv[kMaxViews], orig[kMaxViews];
        int nvalid = 0;
        for (int vi = 0; vi < c.V; vi++) {
            float v = view_cost<NT, N1, PXF, U8, PIN>(c, vi, x, y, pl, ws, rs);
            if (v < kMaxCost) nvalid++;
            else v = kMaxCost;
            cv[vi] = v; orig[vi] = v;
        }
        for (int i = 1; i < c.V; i++) {
            float tmp = cv[i];
            int j = i;
            for (; j >= 1 && tmp < cv[j - 1]; j--) cv[j] = cv[j - 1];
            cv[j] = tmp;
        }
        int nb = nvalid;
        if (c.cost_comb == 1) nb = min(nb, c.n_best);
        if (nb > 0) {
            float s = 0.f;
            for (int i = 0; i < nb; i++) s = fadd(s, cv[i]);
            out.cost = fdiv(s, (float)nb);
            out.ratio = fdiv(cv[0], cv[1]);
            out.beview = -1;
            for (int i = 0; i < c.V; i++) if (cv[0] == orig[i]) out.beview = c.view_id[i];
        } else {
            out.cost = kMaxCost; out.ratio = 0.f; out.beview = -1;
        }
        return out;
    }
}

// ---------------------------------------------------------------------------------------------
// small geometric helpers on the reference camera
// ---------------------------------------------------------------------------------------------
// getDepthFromPlane3_cu / getDisparity_cu (gipuma.cu:436-453): depth of plane (n,d) along the ray of (x,y)
__device__ __forceinline__ float plane_depth(const PmConst &c, const float4 &pl, int x, int y) {
    if (pl.w != pl.w) return 1000.0f;
    const float dy = fsub((float)y, c.cy), dx = fsub((float)x, c.cx);
    const float den = ffma(pl.z, c.fx, ffma(pl.x, dx, fmul(c.alpha, fmul(pl.y, dy))));
    return fdiv(fmul(pl.w, -c.fx), den);
}

// getD_cu (gipuma.cu:71-86): plane offset d so that the plane with normal n passes through the
// point at `depth` on the ray of (x,y)
__device__ __forceinline__ float plane_d(const PmConst &c, float nx, float ny, float nz, int x, int y, float depth) {
    const float ptx = ffma((float)x, depth, -c.Pc[0]);
    const float pty = ffma((float)y, depth, -c.Pc[1]);
    const float ptz = fsub(depth, c.Pc[2]);
    float X, Y, Z;
    matvec3(c.Minv, ptx, pty, ptz, X, Y, Z);
    return -dot3(nx, X, ny, Y, nz, Z);
}

// getViewVector_cu (gipuma.cu:97-105)
__device__ __forceinline__ void view_vector(const PmConst &c, int x, int y, float &vx, float &vy, float &vz) {
    const float ptx = fsub((float)x, c.Pc[0]), pty = fsub((float)y, c.Pc[1]), ptz = fsub(1.0f, c.Pc[2]);
    float X, Y, Z;
    matvec3(c.Minv, ptx, pty, ptz, X, Y, Z);
    vx = fsub(X, c.C[0]); vy = fsub(Y, c.C[1]); vz = fsub(Z, c.C[2]);
    const float r = rsqrtf(dot3(vx, vx, vy, vy, vz, vz));  // normalize_cu, gipuma.cu:88-95
    vx = fmul(vx, r); vy = fmul(vy, r); vz = fmul(vz, r);
}

// curand_uniform on a raw XORWOW output (curand_uniform.h:69-72): x*2^-32 + 2^-33, fused
__device__ __forceinline__ float uniform01(uint32_t u) {
    return ffma(__uint2float_rn(u), __uint_as_float(0x2f800000u), __uint_as_float(0x2f000000u));
}

}  // namespace tsar
