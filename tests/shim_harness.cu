// tests/shim_harness.cu -- TEST INFRASTRUCTURE: plays the role of the reference's runGipuma (main.cpp:1268-1866).
// Builds a managed-memory GlobalState with the layout of include/tsar_gipuma_abi.h exactly as the reference's
// host code would (cudaMallocManaged objects, float textures in cudaArrays), then drives the four drop-in
// entry points firstcuda / sliccuda / fakecuda / fillcuda of libtsar_b200.so and hands the arrays back.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../include/tsar_b200.h"
#include "../include/tsar_gipuma_abi.h"

using namespace tsar_abi;

template <typename T>
static T *managed(size_t n) {
    T *p = nullptr;
    cudaMallocManaged(&p, n * sizeof(T));
    memset(p, 0, n * sizeof(T));
    return p;
}

// reliable flags (lines->scale) for the next run; null = all zero as LineState::resize leaves them
static const float *g_reliable = nullptr;
extern "C" void shim_harness_set_reliable(const float *flags) { g_reliable = flags; }

extern "C" int shim_harness_run(int W, int H, int n_images, const float *const *images, const tsar_camera *cams, float cam_f,
                                const int *subset, int V, const tsar_params *p, const float *canny, int n_regions,
                                const float *region_text, const float *region_norm4, int shipped_flow,
                                const float *import_norm4, const float *import_disp, float *out_norm4, float *out_confid,
                                float *out_fakedepth, float *out_scale) {
    const size_t n = (size_t)W * H;
    tsar_abi::GlobalState *gs = managed<tsar_abi::GlobalState>(1);
    gs->cameras = managed<CameraParameters_cu>(1);
    gs->lines = managed<LineState>(1);
    gs->cannylines = managed<LineState>(1);
    gs->params = managed<AlgorithmParameters>(1);
    gs->col = W; gs->row = H;
    CameraParameters_cu &cp = *gs->cameras;
    cp.f = cam_f; cp.cols = W; cp.rows = H;
    cp.viewSelectionSubset = managed<int>(kMaxImages);
    for (int i = 0; i < V; i++) cp.viewSelectionSubset[i] = subset[i];
    cp.viewSelectionSubsetNumber = V;
    for (int i = 0; i < n_images; i++) {
        Camera_cu &c = cp.cameras[i];
        float **mats[] = {&c.P, &c.P_inv, &c.M_inv, &c.K, &c.K_inv, &c.R, &c.R_orig, &c.R_orig_inv};
        for (float **m : mats) *m = managed<float>(16);
        for (int k = 0; k < 9; k++) {
            c.K[k] = cams[i].K[k]; c.K_inv[k] = cams[i].K_inv[k]; c.R[k] = cams[i].R[k]; c.R_orig[k] = cams[i].R_orig[k];
            c.R_orig_inv[k] = cams[i].R_orig_inv[k]; c.M_inv[k] = cams[i].M_inv[k];
        }
        c.t4 = make_float4(cams[i].t4[0], cams[i].t4[1], cams[i].t4[2], 0);
        c.P_col34 = make_float4(cams[i].P_col34[0], cams[i].P_col34[1], cams[i].P_col34[2], 0);
        c.C4 = make_float4(cams[i].C4[0], cams[i].C4[1], cams[i].C4[2], 0);
        c.fx = cams[i].fx; c.fy = cams[i].fy; c.f = cams[i].f; c.alpha = cams[i].alpha; c.baseline = cams[i].baseline;
        c.depthMin = cams[i].depthMin; c.depthMax = cams[i].depthMax;
        if (p->color_processing) {  // addImageToTextureFloatColor (main.cpp:1150-1188): BGRA float4, the image in channel x
            cudaChannelFormatDesc cd = cudaCreateChannelDesc<float4>();
            cudaMallocArray(&gs->cuArray[i], &cd, W, H);
            std::vector<float4> px(n);
            for (size_t k = 0; k < n; k++) px[k] = make_float4(images[i][k], 7.0f, 255.0f - images[i][k], 0.0f);
            cudaMemcpy2DToArray(gs->cuArray[i], 0, 0, px.data(), (size_t)W * 16, (size_t)W * 16, H, cudaMemcpyHostToDevice);
        } else {
            cudaChannelFormatDesc cd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
            cudaMallocArray(&gs->cuArray[i], &cd, W, H);
            cudaMemcpy2DToArray(gs->cuArray[i], 0, 0, images[i], (size_t)W * 4, (size_t)W * 4, H, cudaMemcpyHostToDevice);
        }
        cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray; rd.res.array.array = gs->cuArray[i];
        cudaTextureDesc td; memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
        td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
        cudaCreateTextureObject(&gs->imgs[i], &rd, &td, NULL);
    }
    AlgorithmParameters &ap = *gs->params;
    ap.box_hsize = p->box_hsize; ap.box_vsize = p->box_vsize; ap.iterations = p->iterations; ap.n_best = p->n_best;
    ap.cost_comb = p->cost_comb; ap.min_disparity = p->min_disparity; ap.max_disparity = p->max_disparity;
    ap.color_processing = p->color_processing != 0; ap.cols = W; ap.rows = H;
    LineState &l = *gs->lines;  // LineState::resize (linestate.h:73-109)
    l.c = managed<float>(n); l.depth = managed<float>(n); l.fakedepth = managed<float>(n); l.norm4 = managed<float4>(n);
    l.scale = managed<float>(n); l.ransa = managed<float>(n); l.canny = managed<float>(n); l.ratio = managed<float>(n);
    l.beview = managed<int>(n); l.lrdiff = managed<float>(n); l.confid = managed<float>(n);
    memcpy(l.canny, canny, n * 4);
    gs->cannylines->text = managed<float>(n_regions);        // Cannyresize (linestate.h:170-190)
    gs->cannylines->norm4 = managed<float4>(n_regions);
    memcpy(gs->cannylines->text, region_text, n_regions * 4);
    memcpy(gs->cannylines->norm4, region_norm4, n_regions * 16);
    if (shipped_flow) {  // main.cpp:1476-1488
        for (size_t i = 0; i < n; i++) {
            l.norm4[i] = make_float4(import_norm4[4 * i], import_norm4[4 * i + 1], import_norm4[4 * i + 2], import_norm4[4 * i + 3]);
            l.c[i] = 1.0f;
            l.depth[i] = import_disp[i];
        }
    }
    ::GlobalState &ref = reinterpret_cast<::GlobalState &>(*gs);
    int rc;
    if ((rc = firstcuda(ref))) return rc;
    if (g_reliable) memcpy(l.scale, g_reliable, n * 4);      // the reliable flags main() reads from weak.png (main.cpp:1499-1514)
    if ((rc = sliccuda(ref))) return rc;
    memcpy(out_confid, l.confid, n * 4);
    if ((rc = fakecuda(ref))) return rc;
    if ((rc = fillcuda(ref))) return rc;
    cudaDeviceSynchronize();
    memcpy(out_norm4, l.norm4, n * 16);
    memcpy(out_fakedepth, l.fakedepth, n * 4);
    memcpy(out_scale, l.scale, n * 4);
    return 0;
}
