// tests/gslicr_harness.cu -- TEST INFRASTRUCTURE.  Plays the reference's gslic() (main.cpp:598-660) against the gSLICr
// drop-in class exported by libtsar_b200.so: this file is compiled against the REFERENCE's own headers
// (gSLICr_Lib/gSLICr.h, globalstate.h -- by oracle/build_ref.sh, where the checkout exists) and linked against
// libtsar_b200.so instead of the reference's gSLICr engine objects.  If it links and runs, main.cpp's call site does.
#include <cuda_runtime.h>
#include <string.h>

#include "gSLICr_Lib/gSLICr.h"

extern "C" int gslicr_harness_run(const unsigned char *rgbx, int w, int h, int spixel_size, int no_iters, float coh_weight,
                                  int enforce, int *labels_out, unsigned char *drawn_out, const char *pgm_path) {
    gSLICr::objects::settings my_settings;                      // main.cpp:608-615
    my_settings.img_size.x = w;
    my_settings.img_size.y = h;
    my_settings.no_segs = 4256;
    my_settings.spixel_size = spixel_size;
    my_settings.coh_weight = coh_weight;
    my_settings.no_iters = no_iters;
    my_settings.color_space = gSLICr::CIELAB;
    my_settings.seg_method = gSLICr::GIVEN_SIZE;
    my_settings.do_enforce_connectivity = enforce != 0;
    gSLICr::engines::core_engine *engine = new gSLICr::engines::core_engine(my_settings);   // main.cpp:633-634
    gSLICr::UChar4Image *in_img = new gSLICr::UChar4Image(my_settings.img_size, true, true);
    gSLICr::UChar4Image *out_img = new gSLICr::UChar4Image(my_settings.img_size, true, true);
    memcpy(in_img->GetData(MEMORYDEVICE_CPU), rgbx, (size_t)w * h * 4);                    // load_image, main.cpp:190-201
    memset(out_img->GetData(MEMORYDEVICE_CPU), 7, (size_t)w * h * 4);
    engine->Process_Frame(in_img, (GlobalState *)0);                                       // main.cpp:646
    const gSLICr::IntImage *seg = engine->Get_Seg_Res();
    int rc = 0;
    if (seg->noDims.x != w || seg->noDims.y != h) rc = 1;
    else memcpy(labels_out, seg->GetData(MEMORYDEVICE_CPU), (size_t)w * h * sizeof(int));
    engine->Draw_Segmentation_Result(out_img);                                             // main.cpp:651
    memcpy(drawn_out, out_img->GetData(MEMORYDEVICE_CPU), (size_t)w * h * 4);
    if (pgm_path) engine->Write_Seg_Res_To_PGM(pgm_path);
    delete engine;
    delete in_img;
    delete out_img;
    return rc;
}
