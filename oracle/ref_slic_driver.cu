// oracle/ref_slic_driver.cu -- TEST INFRASTRUCTURE, not product code.
//
// Builds the reference's gSLICr GPU engine (gSLICr_Lib/engines/gSLICr_seg_engine_GPU.cu, unmodified,
// included from the read-only reference checkout) into oracle/_ref/libgslic_ref.so.  The host file
// gSLICr_seg_engine.cpp does not compile with g++/nvcc (max(int,size_t), lines 143-144, MSVC-only), so
// the three host members it defines are restated here: constructor, destructor and the 12-line kernel
// sequence of Perform_Segmentation (gSLICr_seg_engine.cpp:30-46).  Its CPU adjacency post-pass
// (lines 47-149) computes a result that is discarded and indexes out of bounds: omitted.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "gSLICr_Lib/engines/gSLICr_seg_engine_GPU.cu"

using namespace gSLICr;
using namespace gSLICr::objects;
using namespace gSLICr::engines;

seg_engine::seg_engine(const objects::settings &in_settings) { gSLICr_settings = in_settings; }

seg_engine::~seg_engine() {
    if (source_img != NULL) delete source_img;
    if (cvt_img != NULL) delete cvt_img;
    if (idx_img != NULL) delete idx_img;
    if (spixel_map != NULL) delete spixel_map;
}

void seg_engine::Perform_Segmentation(UChar4Image *in_img, GlobalState *) {
    source_img->SetFrom(in_img, ORUtils::MemoryBlock<Vector4u>::CPU_TO_CUDA);
    Cvt_Img_Space(source_img, cvt_img, gSLICr_settings.color_space);
    Init_Cluster_Centers();
    Find_Center_Association();
    for (int i = 0; i < gSLICr_settings.no_iters; i++) {
        Update_Cluster_Center();
        Find_Center_Association();
    }
    if (gSLICr_settings.do_enforce_connectivity) Enforce_Connectivity();
    cudaDeviceSynchronize();
}

// superpixel records (centre x, y, colour x, y, z, w, id, no_pixels = 8 words each) and the Lab image of the last ref_slic
// call, for stage-by-stage comparisons
namespace {
struct PeekEngine : public seg_engine_GPU {
    PeekEngine(const settings &s) : seg_engine_GPU(s) {}
    SpixelMap *map() { return spixel_map; }
    Float4Image *lab() { return cvt_img; }
};
float *g_centres = nullptr;
int g_n_centres = 0;
float *g_lab = nullptr;
size_t g_n_lab = 0;
}  // namespace

extern "C" long long ref_slic_last_lab(float *out, long long max_pixels) {   // Lab image (4 floats per pixel) of the last call
    const size_t n = g_n_lab < (size_t)max_pixels ? g_n_lab : (size_t)max_pixels;
    if (out && g_lab) memcpy(out, g_lab, n * 16);
    return (long long)g_n_lab;
}

extern "C" int ref_slic_last_centres(float *out, int max_records) {
    const int n = g_n_centres < max_records ? g_n_centres : max_records;
    if (out && g_centres) memcpy(out, g_centres, (size_t)n * 32);
    return g_n_centres;
}

extern "C" int ref_slic(const unsigned char *bgrx, int w, int h, int spixel_size, int no_iters, float coh_weight,
                        int enforce, int *labels_out, float *ms_out) {
    settings s;
    s.img_size.x = w;
    s.img_size.y = h;
    s.no_segs = 4256;            // main.cpp:609 (unused with GIVEN_SIZE)
    s.spixel_size = spixel_size; // main.cpp:610
    s.coh_weight = coh_weight;   // :611
    s.no_iters = no_iters;       // :612
    s.color_space = CIELAB;      // :613
    s.seg_method = GIVEN_SIZE;   // :614
    s.do_enforce_connectivity = enforce != 0;  // :615
    PeekEngine *eng = new PeekEngine(s);
    UChar4Image *in_img = new UChar4Image(s.img_size, true, true);
    memcpy(in_img->GetData(MEMORYDEVICE_CPU), bgrx, (size_t)w * h * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    eng->Perform_Segmentation(in_img, NULL);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms;
    const IntImage *res = eng->Get_Seg_Mask();
    memcpy(labels_out, res->GetData(MEMORYDEVICE_CPU), (size_t)w * h * sizeof(int));
    {
        SpixelMap *m = eng->map();
        m->UpdateHostFromDevice();
        g_n_centres = m->noDims.x * m->noDims.y;
        delete[] g_centres;
        g_centres = new float[(size_t)g_n_centres * 8];
        memcpy(g_centres, m->GetData(MEMORYDEVICE_CPU), (size_t)g_n_centres * 32);
        Float4Image *l = eng->lab();
        l->UpdateHostFromDevice();
        g_n_lab = (size_t)w * h;
        delete[] g_lab;
        g_lab = new float[g_n_lab * 4];
        memcpy(g_lab, l->GetData(MEMORYDEVICE_CPU), g_n_lab * 16);
    }
    cudaError_t err = cudaGetLastError();
    delete in_img;
    delete eng;
    return err == cudaSuccess ? 0 : -3;
}
