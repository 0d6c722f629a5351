/*
 * tsar_gipuma_abi.h -- the reference's C++ entry points, re-hosted on libtsar_b200.so.
 *
 * The reference's boundary for the depthmap path is four free functions taking the managed-memory root
 * object by reference (gipuma.h:2-5):
 *     int firstcuda(GlobalState&); int sliccuda(GlobalState&); int fakecuda(GlobalState&); int fillcuda(GlobalState&);
 * This header declares layout-identical state types so that a caller built against the reference's headers
 * (main.cpp) can link against libtsar_b200.so instead of gipuma.cu without recompiling its own translation
 * units.  Only the LAYOUT is mirrored (field order, types, alignment); the storage strategy is not: the
 * shims copy what a call needs out of the caller's managed memory into the device-resident context of
 * include/tsar_b200.h, run the hand-written kernels, and write the results back.
 * tests/test_cpu.py::test_abi_layout_matches_reference compiles this header next to the reference's own
 * headers (when the reference checkout is present) and compares every offsetof / sizeof.
 *
 * Layout sources: globalstate.h:25-54, linestate.h:10-50, cameraparameters.h:7-27, camera.h:7-29,
 * algorithmparameters.h:54-89, managed.h:5-16 (an empty base with operator new -> cudaMallocManaged).
 */
#ifndef TSAR_GIPUMA_ABI_H
#define TSAR_GIPUMA_ABI_H

#include <cuda_runtime.h>
#include <vector_types.h>

namespace tsar_abi {

constexpr int kMaxImages = 512; /* config.h:2 MAX_IMAGES */

struct Camera_cu { /* camera.h:7-29 ; 176 bytes */
    float *P;
    float4 P_col34;
    float *P_inv, *M_inv, *R, *R_orig, *R_orig_inv;
    float4 t4, C4;
    float fx, fy, f, alpha, baseline;
    bool reference;
    float depthMin, depthMax;
    char *id;
    float *K, *K_inv;
};

struct alignas(128) CameraParameters_cu { /* cameraparameters.h:7-27 */
    float f;
    bool rectified;
    Camera_cu cameras[kMaxImages];
    int idRef, cols, rows;
    int *viewSelectionSubset;
    int viewSelectionSubsetNumber;
};

struct alignas(128) LineState { /* linestate.h:10-50 (pointer table only; the arrays are the caller's) */
    float4 *norm4;
    float *c, *depth, *fakedepth;
    float4 *resize4;
    float *canny;
    int *cenxi, *cenyi, *nein;
    int **neip;
    int *nump;
    int **eacp, **ranp;
    int *pind, *borlen;
    float *depdif, *scale, *ransa, *text;
    float3 *XYZ;
    float *ratio;
    int *beview;
    float *lrdiff, *confid, *ranumax, *size;
    int n, s, l;
};

struct AlgorithmParameters { /* algorithmparameters.h:54-89 */
    int algorithm;
    float max_disparity, min_disparity;
    int box_hsize, box_vsize;
    float tau_color, tau_gradient, alpha, gamma;
    int border_value, iterations;
    bool color_processing;
    float dispTol, normTol, census_epsilon;
    int self_similarity_n;
    float cam_scale;
    int num_img_processed;
    float costThresh, good_factor;
    int n_best, cost_comb;
    bool viewSelection;
    float depthMin, depthMax, min_angle, max_angle, no_texture_sim, no_texture_per;
    unsigned int max_views;
    int cols, rows;
    float thres;
};

struct GlobalState { /* globalstate.h:25-54 */
    CameraParameters_cu *cameras;
    LineState *lines;
    LineState *cannylines;
    void *cs; /* curandState*: allocated by the reference's firstcuda (gipuma.cu:1714); unused here */
    AlgorithmParameters *params;
    int col, row;
    cudaTextureObject_t imgs[kMaxImages];
    cudaArray *cuArray[kMaxImages];
};

static_assert(sizeof(Camera_cu) == 176, "Camera_cu layout (stride read from the reference build: cameras[i] at +16 + 176*i)");
static_assert(offsetof(Camera_cu, t4) == 80 && offsetof(Camera_cu, K) == 152 && offsetof(Camera_cu, K_inv) == 160, "Camera_cu");
static_assert(offsetof(CameraParameters_cu, cameras) == 16, "CameraParameters_cu");
static_assert(offsetof(CameraParameters_cu, cols) == 16 + 176 * kMaxImages + 4, "CameraParameters_cu tail (cols at 90132 in the reference build)");
static_assert(offsetof(CameraParameters_cu, viewSelectionSubset) == 90144 && offsetof(CameraParameters_cu, viewSelectionSubsetNumber) == 90152, "CameraParameters_cu tail");
static_assert(offsetof(GlobalState, imgs) == 48, "GlobalState (imgs at +48 in the reference build)");
static_assert(offsetof(AlgorithmParameters, box_hsize) == 12 && offsetof(AlgorithmParameters, n_best) == 80 && offsetof(AlgorithmParameters, cost_comb) == 84, "AlgorithmParameters");

}  // namespace tsar_abi

/*
 * Drop-in entry points (global namespace, C++ linkage, exactly the reference's mangled names when the caller
 * includes the reference's own gipuma.h: the parameter type is named GlobalState there as well).
 *
 * What each call does with libtsar_b200.so:
 *   firstcuda : env TSAR_B200_PATCHMATCH=1 -> the north-star sequence the reference has commented out
 *               (gipuma.cu:1741-1758: init, iterations x red/black propagation+refinement, L/R check) from the
 *               images/cameras/parameters in gs; seed from TSAR_B200_SEED (default 20240601).
 *               otherwise -> the shipped behaviour: gipuma_get_disp on the imported normals/disparities
 *               (gipuma.cu:1755).  Writes lines->norm4 (+ c, ratio, beview, lrdiff in PatchMatch mode).
 *   sliccuda  : gipuma_getview (gipuma.cu:1806): lines->confid, lines->depth.
 *   fakecuda  : gipuma_update_scale_2 (gipuma.cu:1875): lines->fakedepth.
 *   fillcuda  : gipuma_update_scale + gipuma_compute_disp (gipuma.cu:1842-1848): lines->norm4, c, scale, depth.
 * All return 0 like the reference; on a CUDA/library error they print the message and return a negative
 * tsar_status instead of calling exit() (helper_cuda.h:891-905 does).
 */
struct GlobalState;
int firstcuda(GlobalState &gs);
int sliccuda(GlobalState &gs);
int fakecuda(GlobalState &gs);
int fillcuda(GlobalState &gs);

#endif /* TSAR_GIPUMA_ABI_H */
