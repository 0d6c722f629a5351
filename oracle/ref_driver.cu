// oracle/ref_driver.cu -- TEST INFRASTRUCTURE, not product code.
//
// Builds the reference's own depthmap kernels (gipuma.cu, unmodified, included from the read-only
// reference checkout via -I) into oracle/_ref/libtsar_ref*.so and exposes them behind a tiny C ABI
// so tests/, __graft_entry__.smoke() and bench.py --impl reference can run the *reference's*
// arithmetic on the same inputs as the product library.  Nothing in the product path links or
// loads this file.
//
// Interventions (all at driver / build-line level, the device code that produces the numbers is
// the reference's):
//   (i)   `#define clock64() g_oracle_seed` after the CUDA headers and before the include, so that
//         curand_init(clock64(), y, x) (gipuma.cu:700,1077) is reproducible;
//   (ii)  lines->c is allocated with a zero-filled guard in front (out-of-bounds read of
//         c[pindex - 3*cols] for y < 3, gipuma.cu:906);
//   (iii) the launch sequence of the commented block gipuma.cu:1741-1754 is re-created here;
//   (iv)  cameras are filled from the caller's values instead of OpenCV (cameraGeometryUtils.h);
//   (v)   -DORACLE_SNAPSHOT build only: oracle/build_ref.sh compiles a temp copy of gipuma.cu in
//         which the two final stores of gipuma_checkerboard_spatialProp_cu (gipuma.cu:1047-1048)
//         go to lines->ransa / lines->resize4; the driver then copies them back for the launch's
//         own colour.  This removes the same-colour read/write race and gives a
//         deterministic oracle with Jacobi (pre-launch snapshot) semantics.  The same is done for the
//         own-pixel write-back blocks of gipuma_WMF (gipuma.cu:1686-1695 -> ransa) and gipuma_WMF_Final
//         (gipuma.cu:1470-1485 -> resize4 / fakedepth / ransa), which race on scale / depth / norm4.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

__device__ unsigned long long g_oracle_seed;
#define clock64() g_oracle_seed
#ifdef ORACLE_SNAPSHOT
#include "gipuma_snapshot.cu" /* generated into a temp dir by oracle/build_ref.sh, never committed */
#else
#include "gipuma.cu"
#endif
#undef clock64

#include "tsar_b200.h" /* tsar_camera / tsar_params / tsar_field only (shared plain structs) */

#define RCHECK(x)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "[ref] %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            return -3;                                                                     \
        }                                                                                  \
    } while (0)

struct RefCtx {
    GlobalState *gs = nullptr;
    AlgorithmParameters *params = nullptr;
    int W = 0, H = 0, n_images = 0, n_regions = 0;
    float *c_base = nullptr;    // guard-padded allocation backing lines->c
    float *c_resize_orig = nullptr;
    size_t guard = 0;
    dim3 grid_cb, block_cb, grid_px, block_px;
    bool colour = false;  // color_processing: BGRA float4 textures, kernels instantiated for float4 (gipuma.cu:1881-1912)
};

// wrapper kernel: the reference's pmCostMultiview_cu on explicit (pixel, plane) pairs
__global__ void ref_eval_kernel(GlobalState &gs, int n, const int2 *xy, const float4 *planes, float *cost,
                                int *bv, float *ratio) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int hr = (gs.params->box_hsize - 1) / 2, vr = (gs.params->box_vsize - 1) / 2;
    int b = -2;
    float r = 0.f;
    cost[i] = pmCostMultiview_cu<float>(b, r, gs.imgs, xy[i], planes[i], vr, hr, *gs.params, *gs.cameras,
                                        gs.lines->norm4, 0);
    bv[i] = b;
    ratio[i] = r;
}

#ifdef ORACLE_SNAPSHOT
// y_limit: rows reached by the reference's checkerboard grid (gipuma.cu:1721; for some odd heights the last
// row is never processed, so there is nothing to merge there)
__global__ void ref_merge_colour(GlobalState &gs, int colour, int y_limit) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= gs.cameras->cols || y >= gs.cameras->rows || y >= y_limit) return;
    if (((x + y) & 1) != colour) return;
    int p = y * gs.cameras->cols + x;
    gs.lines->c[p] = gs.lines->ransa[p];
    gs.lines->norm4[p] = gs.lines->resize4[p];
}
#endif

static void fill_cam(Camera_cu &c, const tsar_camera &s) {
    memset(c.P, 0, 16 * sizeof(float));
    memset(c.P_inv, 0, 16 * sizeof(float));
    for (int k = 0; k < 9; k++) {
        c.K[k] = s.K[k];
        c.K_inv[k] = s.K_inv[k];
        c.R[k] = s.R[k];
        c.R_orig[k] = s.R_orig[k];
        c.R_orig_inv[k] = s.R_orig_inv[k];
        c.M_inv[k] = s.M_inv[k];
    }
    c.t4 = make_float4(s.t4[0], s.t4[1], s.t4[2], 0.f);
    c.P_col34 = make_float4(s.P_col34[0], s.P_col34[1], s.P_col34[2], 0.f);
    c.C4 = make_float4(s.C4[0], s.C4[1], s.C4[2], 0.f);
    c.fx = s.fx;
    c.fy = s.fy;
    c.f = s.f;
    c.alpha = s.alpha;
    c.baseline = s.baseline;
    c.depthMin = s.depthMin;
    c.depthMax = s.depthMax;
}

extern "C" {

const char *ref_variant(void) {
#if defined(ORACLE_SNAPSHOT) && defined(ORACLE_WMF_INIT)
    return "snapshot_init";   // snapshot twin + zero-initialised norm_mid in gipuma_WMF / gipuma_WMF_Final (build_ref.sh 2b)
#elif defined(ORACLE_SNAPSHOT)
    return "snapshot";
#else
    return "asis";
#endif
}

int ref_create(int W, int H, int n_images, const float *const *images, const tsar_camera *cams, float cam_f,
               const int *subset, int V, const tsar_params *p, void **out) {
    if (!out || W <= 0 || H <= 0 || n_images < 1 || n_images > MAX_IMAGES || V < 0 || V > 32) return -1;
    RefCtx *r = new RefCtx;
    r->W = W;
    r->H = H;
    r->n_images = n_images;
    r->gs = new GlobalState;  // managed, as main.cpp:1876
    r->params = new AlgorithmParameters;
    GlobalState &gs = *r->gs;
    // LineState has no constructor: its pointer members hold whatever the managed block held before, and
    // ~LineState cudaFree()s all of them.  The reference program builds one GlobalState per process and
    // resize()s both states; a test process builds many, so start both from null pointers.
    memset(gs.lines, 0, sizeof(LineState));
    memset(gs.cannylines, 0, sizeof(LineState));
    AlgorithmParameters &ap = *r->params;
    ap.box_hsize = p->box_hsize;
    ap.box_vsize = p->box_vsize;
    ap.iterations = p->iterations;
    ap.n_best = p->n_best;
    ap.cost_comb = p->cost_comb;
    ap.min_disparity = p->min_disparity;
    ap.max_disparity = p->max_disparity;
    ap.color_processing = p->color_processing != 0;
    r->colour = ap.color_processing;
#ifndef ORACLE_SNAPSHOT
    if (r->colour) return -1;  // the float4 instantiations are only built into the snapshot variant
#endif
    ap.cols = W;
    ap.rows = H;
    ap.depthMin = cams[0].depthMin;
    ap.depthMax = cams[0].depthMax;
    gs.params = &ap;  // main.cpp:1408
    gs.col = W;       // main.cpp:624-625
    gs.row = H;
    gs.cameras->cols = W;
    gs.cameras->rows = H;
    gs.cameras->f = cam_f;
    for (int i = 0; i < n_images; i++) fill_cam(gs.cameras->cameras[i], cams[i]);
    gs.cameras->cameras[0].reference = true;
    for (int i = 0; i < V; i++) gs.cameras->viewSelectionSubset[i] = subset[i];
    gs.cameras->viewSelectionSubsetNumber = V;

    size_t n = (size_t)W * H;
    gs.lines->resize((int)n);  // main.cpp:554
    // (ii) guard-padded c[]
    r->guard = (size_t)4 * W;
    r->c_resize_orig = gs.lines->c;
    RCHECK(cudaMallocManaged(&r->c_base, (n + r->guard) * sizeof(float)));
    memset(r->c_base, 0, (n + r->guard) * sizeof(float));
    gs.lines->c = r->c_base + r->guard;
    RCHECK(cudaMallocManaged(&gs.lines->text, n * sizeof(float)));
    memset(gs.lines->text, 0, n * sizeof(float));
#ifdef ORACLE_SNAPSHOT
    RCHECK(cudaMallocManaged(&gs.lines->resize4, n * sizeof(float4)));
    memset(gs.lines->resize4, 0, n * sizeof(float4));
#endif
    // textures exactly as addImageToTextureFloatGray (main.cpp:1190-1228) / addImageToTextureFloatColor (1150-1188);
    // in colour mode channel x carries the image, the other channels carry different data (they must not matter)
    for (int i = 0; i < n_images; i++) {
        if (r->colour) {
            cudaChannelFormatDesc cd = cudaCreateChannelDesc<float4>();
            RCHECK(cudaMallocArray(&gs.cuArray[i], &cd, W, H));
            std::vector<float4> px(n);
            for (size_t k = 0; k < n; k++) px[k] = make_float4(images[i][k], 0.5f * images[i][k] + 3.0f, 255.0f - images[i][k], 0.0f);
            RCHECK(cudaMemcpy2DToArray(gs.cuArray[i], 0, 0, px.data(), (size_t)W * sizeof(float4), (size_t)W * sizeof(float4), H,
                                       cudaMemcpyHostToDevice));
        } else {
            cudaChannelFormatDesc cd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
            RCHECK(cudaMallocArray(&gs.cuArray[i], &cd, W, H));
            RCHECK(cudaMemcpy2DToArray(gs.cuArray[i], 0, 0, images[i], (size_t)W * sizeof(float),
                                       (size_t)W * sizeof(float), H, cudaMemcpyHostToDevice));
        }
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = gs.cuArray[i];
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = cudaAddressModeWrap;
        td.addressMode[1] = cudaAddressModeWrap;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        RCHECK(cudaCreateTextureObject(&gs.imgs[i], &rd, &td, NULL));
    }
    RCHECK(cudaMalloc(&gs.cs, n * sizeof(curandState)));  // gipuma.cu:1714
    RCHECK(cudaMemset(gs.cs, 0, n * sizeof(curandState)));
    // launch geometry of gipuma_first (gipuma.cu:1716-1731)
    r->block_cb = dim3(32, 16);
    r->grid_cb = dim3((W + 31) / 32, ((H / 2) + 15) / 16);
    r->block_px = dim3(16, 16);
    r->grid_px = dim3((W + 15) / 16, (H + 15) / 16);
    cudaDeviceSetCacheConfig(cudaFuncCachePreferShared);  // gipuma.cu:1703
    *out = r;
    return 0;
}

int ref_destroy(void *h) {
    RefCtx *r = (RefCtx *)h;
    if (!r) return 0;
    cudaDeviceSynchronize();
    for (int i = 0; i < r->n_images; i++) {
        cudaDestroyTextureObject(r->gs->imgs[i]);
        cudaFreeArray(r->gs->cuArray[i]);
    }
    cudaFree(r->gs->cs);
    r->gs->lines->c = r->c_resize_orig;
    cudaFree(r->c_base);
    delete r->gs;
    delete r->params;
    delete r;
    cudaGetLastError();  // nothing from teardown may surface in the next engine's launch checks
    return 0;
}

int ref_set_regions(void *h, int n_regions, const float *text, const float *norm4) {
    RefCtx *r = (RefCtx *)h;
    r->gs->cannylines->Cannyresize(n_regions);  // main.cpp:571
    r->n_regions = n_regions;
    memcpy(r->gs->cannylines->text, text, n_regions * sizeof(float));
    memcpy(r->gs->cannylines->norm4, norm4, n_regions * sizeof(float4));
    return 0;
}

static void *field_ptr(RefCtx *r, int field, size_t *elt, size_t *count) {
    LineState *l = r->gs->lines;
    size_t n = (size_t)r->W * r->H;
    *count = n;
    *elt = 4;
    switch (field) {
        case TSAR_F_NORM4: *elt = 16; return l->norm4;
        case TSAR_F_COST: return l->c;
        case TSAR_F_DEPTH: return l->depth;
        case TSAR_F_FAKEDEPTH: return l->fakedepth;
        case TSAR_F_SCALE: return l->scale;
        case TSAR_F_CANNY: return l->canny;
        case TSAR_F_RATIO: return l->ratio;
        case TSAR_F_BEVIEW: return l->beview;
        case TSAR_F_LRDIFF: return l->lrdiff;
        case TSAR_F_CONFID: return l->confid;
        case TSAR_F_REGION_TEXT: *count = r->n_regions; return r->gs->cannylines->text;
        case TSAR_F_REGION_NORM4: *count = r->n_regions; *elt = 16; return r->gs->cannylines->norm4;
    }
    return nullptr;
}

int ref_upload(void *h, int field, const void *src, size_t bytes) {
    RefCtx *r = (RefCtx *)h;
    size_t elt, count;
    void *p = field_ptr(r, field, &elt, &count);
    if (!p || bytes != elt * count) return -1;
    RCHECK(cudaDeviceSynchronize());
    RCHECK(cudaMemcpy(p, src, bytes, cudaMemcpyDefault));
    return 0;
}

int ref_download(void *h, int field, void *dst, size_t bytes) {
    RefCtx *r = (RefCtx *)h;
    size_t elt, count;
    void *p = field_ptr(r, field, &elt, &count);
    if (!p || bytes != elt * count) return -1;
    RCHECK(cudaDeviceSynchronize());
    RCHECK(cudaMemcpy(dst, p, bytes, cudaMemcpyDefault));
    return 0;
}

static int set_seed(uint64_t seed) {
    unsigned long long s = seed;
    RCHECK(cudaMemcpyToSymbol(g_oracle_seed, &s, sizeof(s)));
    return 0;
}

// template dispatch on color_processing as gipuma.cu:1881-1912 does
#ifdef ORACLE_SNAPSHOT
#define TLAUNCH(K, grid, block, ...)                                        \
    do {                                                                    \
        if (r->colour) K<float4><<<grid, block>>>(__VA_ARGS__);             \
        else K<float><<<grid, block>>>(__VA_ARGS__);                        \
    } while (0)
#else
#define TLAUNCH(K, grid, block, ...) K<float><<<grid, block>>>(__VA_ARGS__)
#endif

int ref_init(void *h, uint64_t seed) {  // gipuma.cu:1741
    RefCtx *r = (RefCtx *)h;
    if (set_seed(seed)) return -3;
    TLAUNCH(gipuma_init_cu2, r->grid_px, r->block_px, *r->gs);
    RCHECK(cudaGetLastError());
    RCHECK(cudaDeviceSynchronize());
    return 0;
}

static int launch_kind(RefCtx *r, int kind, uint64_t seed, bool sync) {
    GlobalState &gs = *r->gs;
    switch (kind) {
        case TSAR_BLACK_SPATIAL:
            TLAUNCH(gipuma_black_spatialProp_cu, r->grid_cb, r->block_cb, gs, false);
#ifdef ORACLE_SNAPSHOT
            ref_merge_colour<<<r->grid_px, r->block_px>>>(gs, 0, 32 * (int)r->grid_cb.y);
#endif
            break;
        case TSAR_BLACK_REFINE:
            if (set_seed(seed)) return -3;
            TLAUNCH(gipuma_black_planeRefine_cu, r->grid_cb, r->block_cb, gs, false);
            break;
        case TSAR_RED_SPATIAL:
            TLAUNCH(gipuma_red_spatialProp_cu, r->grid_cb, r->block_cb, gs, false);
#ifdef ORACLE_SNAPSHOT
            ref_merge_colour<<<r->grid_px, r->block_px>>>(gs, 1, 32 * (int)r->grid_cb.y);
#endif
            break;
        case TSAR_RED_REFINE:
            if (set_seed(seed)) return -3;
            TLAUNCH(gipuma_red_planeRefine_cu, r->grid_cb, r->block_cb, gs, false);
            break;
        default:
            return -1;
    }
    RCHECK(cudaGetLastError());
    if (sync) RCHECK(cudaDeviceSynchronize());  // gipuma.cu:1747 (a device sync after every kernel)
    return 0;
}

int ref_launch(void *h, int kind, uint64_t seed) { return launch_kind((RefCtx *)h, kind, seed, true); }

int ref_iterate(void *h, int iters, uint64_t seed0) {  // gipuma.cu:1744-1754
    RefCtx *r = (RefCtx *)h;
    for (int it = 0; it < iters; it++) {
        int e;
        if ((e = launch_kind(r, TSAR_BLACK_SPATIAL, 0, true))) return e;
        if ((e = launch_kind(r, TSAR_BLACK_REFINE, seed0 + 1 + 2 * it, true))) return e;
        if ((e = launch_kind(r, TSAR_RED_SPATIAL, 0, true))) return e;
        if ((e = launch_kind(r, TSAR_RED_REFINE, seed0 + 2 + 2 * it, true))) return e;
    }
    return 0;
}

#define PX_KERNEL(name, call)                                \
    int name(void *h) {                                      \
        RefCtx *r = (RefCtx *)h;                             \
        call<<<r->grid_px, r->block_px>>>(*r->gs);           \
        RCHECK(cudaGetLastError());                          \
        RCHECK(cudaDeviceSynchronize());                     \
        return 0;                                            \
    }
int ref_lrdiff(void *h) {  // gipuma.cu:1758 (samples the textures: typed like the propagation kernels)
    RefCtx *r = (RefCtx *)h;
    TLAUNCH(gipuma_getlrdiff, r->grid_px, r->block_px, *r->gs);
    RCHECK(cudaGetLastError());
    RCHECK(cudaDeviceSynchronize());
    return 0;
}
PX_KERNEL(ref_getview, gipuma_getview<float>)              // gipuma.cu:1806
PX_KERNEL(ref_get_disp, gipuma_get_disp<float>)            // gipuma.cu:1755
PX_KERNEL(ref_update_scale_2, gipuma_update_scale_2<float>)  // gipuma.cu:1875
PX_KERNEL(ref_update_scale, gipuma_update_scale<float>)    // gipuma.cu:1842
PX_KERNEL(ref_compute_disp, gipuma_compute_disp)           // gipuma.cu:1848

int ref_wmf(void *h, int iter) {  // gipuma.cu:1810
    RefCtx *r = (RefCtx *)h;
    LineState *l = r->gs->lines;
    const size_t n = (size_t)r->W * r->H;
    (void)l; (void)n;
#ifdef ORACLE_SNAPSHOT
    RCHECK(cudaDeviceSynchronize());
    RCHECK(cudaMemcpy(l->ransa, l->scale, n * 4, cudaMemcpyDefault));  // snapshot variant: results land in ransa
#endif
    gipuma_WMF<float><<<r->grid_px, r->block_px>>>(*r->gs, iter);
    RCHECK(cudaGetLastError());
    RCHECK(cudaDeviceSynchronize());
#ifdef ORACLE_SNAPSHOT
    RCHECK(cudaMemcpy(l->scale, l->ransa, n * 4, cudaMemcpyDefault));
#endif
    return 0;
}
int ref_wmf_final(void *h, int iter) {  // gipuma.cu:1845
    RefCtx *r = (RefCtx *)h;
    LineState *l = r->gs->lines;
    const size_t n = (size_t)r->W * r->H;
    (void)l; (void)n;
#ifdef ORACLE_SNAPSHOT
    RCHECK(cudaDeviceSynchronize());  // snapshot variant: own-pixel results land in resize4 / fakedepth / ransa
    RCHECK(cudaMemcpy(l->resize4, l->norm4, n * 16, cudaMemcpyDefault));
    RCHECK(cudaMemcpy(l->fakedepth, l->depth, n * 4, cudaMemcpyDefault));
    RCHECK(cudaMemcpy(l->ransa, l->scale, n * 4, cudaMemcpyDefault));
#endif
    gipuma_WMF_Final<float><<<r->grid_px, r->block_px>>>(*r->gs, iter);
    RCHECK(cudaGetLastError());
    RCHECK(cudaDeviceSynchronize());
#ifdef ORACLE_SNAPSHOT
    RCHECK(cudaMemcpy(l->norm4, l->resize4, n * 16, cudaMemcpyDefault));
    RCHECK(cudaMemcpy(l->depth, l->fakedepth, n * 4, cudaMemcpyDefault));
    RCHECK(cudaMemcpy(l->scale, l->ransa, n * 4, cudaMemcpyDefault));
#endif
    return 0;
}

int ref_eval_planes(void *h, int n, const int *xy, const float *planes, float *cost, int *beview, float *ratio) {
    RefCtx *r = (RefCtx *)h;
    int2 *dxy;
    float4 *dpl;
    float *dc, *dr;
    int *db;
    RCHECK(cudaMalloc(&dxy, n * sizeof(int2)));
    RCHECK(cudaMalloc(&dpl, n * sizeof(float4)));
    RCHECK(cudaMalloc(&dc, n * sizeof(float)));
    RCHECK(cudaMalloc(&dr, n * sizeof(float)));
    RCHECK(cudaMalloc(&db, n * sizeof(int)));
    RCHECK(cudaMemcpy(dxy, xy, n * sizeof(int2), cudaMemcpyHostToDevice));
    RCHECK(cudaMemcpy(dpl, planes, n * sizeof(float4), cudaMemcpyHostToDevice));
    ref_eval_kernel<<<(n + 127) / 128, 128>>>(*r->gs, n, dxy, dpl, dc, db, dr);
    RCHECK(cudaGetLastError());
    RCHECK(cudaDeviceSynchronize());
    RCHECK(cudaMemcpy(cost, dc, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (beview) RCHECK(cudaMemcpy(beview, db, n * sizeof(int), cudaMemcpyDeviceToHost));
    if (ratio) RCHECK(cudaMemcpy(ratio, dr, n * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(dxy);
    cudaFree(dpl);
    cudaFree(dc);
    cudaFree(dr);
    cudaFree(db);
    return 0;
}

// Whole north-star sequence with the reference's own sync-after-every-kernel structure, timed with
// CUDA events on the legacy default stream (the stream the reference launches on).
// prefetch != 0: cudaMemPrefetchAsync of the managed per-pixel arrays first (BASELINE.md B-ref-GPU).
int ref_depthmap(void *h, uint64_t seed0, int iters, int prefetch, float *ms_out) {
    RefCtx *r = (RefCtx *)h;
    GlobalState &gs = *r->gs;
    size_t n = (size_t)r->W * r->H;
    if (prefetch) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaMemPrefetchAsync(gs.lines->norm4, n * 16, dev, 0);
        cudaMemPrefetchAsync(r->c_base, (n + r->guard) * 4, dev, 0);
        cudaMemPrefetchAsync(gs.lines->ratio, n * 4, dev, 0);
        cudaMemPrefetchAsync(gs.lines->beview, n * 4, dev, 0);
        cudaMemPrefetchAsync(gs.lines->lrdiff, n * 4, dev, 0);
        cudaMemPrefetchAsync(gs.lines->confid, n * 4, dev, 0);
        cudaMemPrefetchAsync(gs.lines->depth, n * 4, dev, 0);
        RCHECK(cudaDeviceSynchronize());
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int e;
    if ((e = ref_init(h, seed0))) return e;
    if ((e = ref_iterate(h, iters, seed0))) return e;
    if ((e = ref_lrdiff(h))) return e;
    if ((e = ref_getview(h))) return e;
    if ((e = ref_compute_disp(h))) return e;
    cudaEventRecord(e1);
    RCHECK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

}  // extern "C"
