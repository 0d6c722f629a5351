// gslicr_shim.cu -- gSLICr::engines::core_engine (gSLICr_Lib/engines/gSLICr_core_engine.h:11-33) on top of tsar_slic, with
// the reference's own mangled names, so that the reference's gslic() (main.cpp:598-660: `new core_engine(settings)`,
// Process_Frame, Draw_Segmentation_Result, delete) links against libtsar_b200.so unchanged.
//
// The caller is compiled against the reference's headers; this file must not define any ORUtils / gSLICr template of its
// own under the reference's names (their weak vtable / inline symbols would collide with the caller's), so the argument
// types are only DECLARED here -- a pointer to an incomplete type mangles like a pointer to the complete one -- and the
// objects the caller hands in are read through layout mirrors (ImageMirror, SettingsMirror) whose offsets
// tests/test_cpu.py::test_gslicr_shim_layout_matches_reference checks against the reference's headers.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <vector>

#include "../../include/tsar_b200.h"
#include "../../include/tsar_gslicr_abi.h"

struct GlobalState;
namespace ORUtils {
template <class T> class Vector4;
template <class T> class Image;
}  // namespace ORUtils

namespace gSLICr {
typedef ORUtils::Image<ORUtils::Vector4<unsigned char>> UChar4Image;
typedef ORUtils::Image<int> IntImage;
namespace objects { struct settings; }
namespace engines {
class seg_engine;
class core_engine {
  private:
    seg_engine *slic_seg_engine;   // the one data member of the reference's class: here it carries the shim's state

  public:
    core_engine(const objects::settings &in_settings);
    ~core_engine();
    void Process_Frame(UChar4Image *in_img, GlobalState *gs);
    const IntImage *Get_Seg_Res();
    void Draw_Segmentation_Result(UChar4Image *out_img);
    void Write_Seg_Res_To_PGM(const char *fileName);
};
}  // namespace engines
}  // namespace gSLICr

namespace {

struct ShimEngine {
    tsar_ctx *ctx = nullptr;
    tsar_slic_settings cfg{};
    std::vector<int> labels;
    std::vector<unsigned char> source;       // last input frame (B, G, R, x), for Draw_Segmentation_Result
    tsar_gslicr_abi::ImageMirror seg_res{};  // what Get_Seg_Res hands out: noDims + data_cpu valid, nothing else
};

ShimEngine *state_of(gSLICr::engines::seg_engine *p) { return reinterpret_cast<ShimEngine *>(p); }

}  // namespace

gSLICr::engines::core_engine::core_engine(const objects::settings &in_settings) {
    const tsar_gslicr_abi::SettingsMirror &s = reinterpret_cast<const tsar_gslicr_abi::SettingsMirror &>(in_settings);
    ShimEngine *e = new ShimEngine;
    e->cfg.img_w = s.img_w;
    e->cfg.img_h = s.img_h;
    if (s.seg_method == 0) {   // GIVEN_NUM: gSLICr_seg_engine_GPU.cu:65-70
        const float cluster = (float)(s.img_w * s.img_h) / (float)s.no_segs;
        e->cfg.spixel_size = (int)ceilf(sqrtf(cluster));
    } else {
        e->cfg.spixel_size = s.spixel_size;
    }
    e->cfg.no_iters = s.no_iters;
    e->cfg.coh_weight = s.coh_weight;
    e->cfg.do_enforce_connectivity = s.do_enforce_connectivity ? 1 : 0;
    e->cfg.correct_reduction = 0;   // parity with the reference build
    if (s.color_space != 0)         // CIELAB = 0 (gSLICr_defines.h:71-76); TSAR's call site uses nothing else (main.cpp:612)
        fprintf(stderr, "[tsar_b200 gSLICr shim] colour space %d requested; only CIELAB is implemented (main.cpp:612)\n", s.color_space);
    int dev = 0;
    cudaGetDevice(&dev);
    if (tsar_create(dev, nullptr, &e->ctx) != TSAR_OK) {
        fprintf(stderr, "[tsar_b200 gSLICr shim] no sm_100 device: there is no CPU path\n");
        exit(EXIT_FAILURE);         // the reference's ORcudaSafeCall exits as well (ORUtils/CUDADefines.h:27-36)
    }
    e->labels.assign((size_t)s.img_w * s.img_h, 0);
    slic_seg_engine = reinterpret_cast<seg_engine *>(e);
}

gSLICr::engines::core_engine::~core_engine() {
    ShimEngine *e = state_of(slic_seg_engine);
    if (!e) return;
    tsar_destroy(e->ctx);
    delete e;
    slic_seg_engine = nullptr;
}

// Perform_Segmentation (gSLICr_seg_engine.cpp:30-46).  `gs` is only read by the reference's CPU adjacency post-pass, whose
// result is discarded: unused here.
void gSLICr::engines::core_engine::Process_Frame(UChar4Image *in_img, GlobalState *) {
    ShimEngine *e = state_of(slic_seg_engine);
    const tsar_gslicr_abi::ImageMirror *im = reinterpret_cast<const tsar_gslicr_abi::ImageMirror *>(in_img);
    if (im->dims_x != e->cfg.img_w || im->dims_y != e->cfg.img_h || !im->data_cpu) {
        fprintf(stderr, "[tsar_b200 gSLICr shim] frame %dx%d does not match the engine's %dx%d\n", im->dims_x, im->dims_y, e->cfg.img_w, e->cfg.img_h);
        exit(EXIT_FAILURE);
    }
    const size_t n = (size_t)e->cfg.img_w * e->cfg.img_h;
    e->source.assign((const unsigned char *)im->data_cpu, (const unsigned char *)im->data_cpu + 4 * n);
    const int rc = tsar_slic(e->ctx, e->source.data(), &e->cfg, e->labels.data());
    if (rc != TSAR_OK) {
        fprintf(stderr, "[tsar_b200 gSLICr shim] tsar_slic failed (%d): %s\n", rc, tsar_last_error(e->ctx));
        exit(EXIT_FAILURE);
    }
}

const gSLICr::IntImage *gSLICr::engines::core_engine::Get_Seg_Res() {
    ShimEngine *e = state_of(slic_seg_engine);
    e->seg_res.isAllocated_CPU = true;
    e->seg_res.data_cpu = e->labels.data();
    e->seg_res.data_cuda = nullptr;
    e->seg_res.dataSize = e->labels.size();
    e->seg_res.dims_x = e->cfg.img_w;
    e->seg_res.dims_y = e->cfg.img_h;
    return reinterpret_cast<const IntImage *>(&e->seg_res);
}

// Draw_Segmentation_Result_device (GPU.cu:223-232, shared.h:136-151): interior pixels whose label differs from a
// 4-neighbour become (0, 0, 255, 0), the others keep the source colour; border pixels are not written.
void gSLICr::engines::core_engine::Draw_Segmentation_Result(UChar4Image *out_img) {
    ShimEngine *e = state_of(slic_seg_engine);
    tsar_gslicr_abi::ImageMirror *om = reinterpret_cast<tsar_gslicr_abi::ImageMirror *>(out_img);
    const int w = e->cfg.img_w, h = e->cfg.img_h;
    if (om->dims_x != w || om->dims_y != h || !om->data_cpu || e->source.empty()) return;
    unsigned char *out = (unsigned char *)om->data_cpu;
    const int *idx = e->labels.data();
    for (int y = 1; y <= h - 2; y++)
        for (int x = 1; x <= w - 2; x++) {
            const int i = y * w + x;
            const bool edge = idx[i] != idx[i + 1] || idx[i] != idx[i - 1] || idx[i] != idx[i - w] || idx[i] != idx[i + w];
            if (edge) { out[4 * i] = 0; out[4 * i + 1] = 0; out[4 * i + 2] = 255; out[4 * i + 3] = 0; }
            else memcpy(out + 4 * i, e->source.data() + 4 * i, 4);
        }
}

// gSLICr_core_engine.cpp:31-45: 16-bit big-endian PGM of the label map
void gSLICr::engines::core_engine::Write_Seg_Res_To_PGM(const char *fileName) {
    ShimEngine *e = state_of(slic_seg_engine);
    const int w = e->cfg.img_w, h = e->cfg.img_h;
    std::ofstream f(fileName, std::ios_base::out | std::ios_base::binary | std::ios_base::trunc);
    f << "P5\n" << w << " " << h << "\n65535\n";
    for (int i = 0; i < w * h; i++) {
        const unsigned short lab = (unsigned short)e->labels[i];
        const unsigned short be = (unsigned short)(lab << 8 | lab >> 8);
        f.write((const char *)&be, sizeof(be));
    }
}
