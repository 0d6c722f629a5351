/*
 * tsar_gslicr_abi.h -- layout mirrors for the gSLICr drop-in class in libtsar_b200.so.
 *
 * The reference's superpixel boundary is a C++ class, gSLICr::engines::core_engine
 * (gSLICr_Lib/engines/gSLICr_core_engine.h:11-33; call site main.cpp:633-651):
 *     core_engine(const objects::settings&);  ~core_engine();
 *     void Process_Frame(UChar4Image*, GlobalState*);  const IntImage* Get_Seg_Res();
 *     void Draw_Segmentation_Result(UChar4Image*);      void Write_Seg_Res_To_PGM(const char*);
 * libtsar_b200.so exports these members under the reference's own mangled names (tsar-mvs_b200/csrc/gslicr_shim.cu),
 * implemented on tsar_slic.  The caller keeps its own ORUtils / gSLICr headers; the shim reads the caller's objects
 * through the mirrors below.  tests/test_cpu.py::test_gslicr_shim_layout_matches_reference compiles them next to the
 * reference's headers and compares every offsetof / sizeof.
 *
 * Layout sources: gSLICr_Lib/objects/gSLICr_settings.h:10-21, ORUtils/MemoryBlock.h:33-56 (a polymorphic class: the
 * vtable pointer comes first), ORUtils/Image.h:16-20.
 */
#ifndef TSAR_GSLICR_ABI_H
#define TSAR_GSLICR_ABI_H

#include <stddef.h>

namespace tsar_gslicr_abi {

struct SettingsMirror { /* gSLICr::objects::settings */
    int img_w, img_h;   /* Vector2i img_size */
    int no_segs;
    int spixel_size;
    int no_iters;
    float coh_weight;
    bool do_enforce_connectivity;
    int color_space;    /* COLOR_SPACE: CIELAB = 0, XYZ, RGB */
    int seg_method;     /* SEG_METHOD: GIVEN_NUM = 0, GIVEN_SIZE */
};

struct ImageMirror { /* ORUtils::Image<T> : MemoryBlock<T> */
    void *vptr;         /* MemoryBlock has a virtual destructor */
    bool isAllocated_CPU, isAllocated_CUDA, isMetalCompatible;
    void *data_cpu;
    void *data_cuda;
    size_t dataSize;
    int dims_x, dims_y; /* Vector2<int> noDims */
};

}  /* namespace tsar_gslicr_abi */
#endif
