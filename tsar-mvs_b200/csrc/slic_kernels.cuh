// slic_kernels.cuh -- gSLICr superpixel segmentation (placeholder until the kernels land)
#pragma once
#include <cuda_runtime.h>
#include "../../include/tsar_b200.h"
namespace tsar {
struct SlicState {};
static inline void slic_free(SlicState &) {}
static inline const char *slic_run(SlicState &, const unsigned char *, const tsar_slic_settings &, int *, cudaStream_t, int *nl) {
    *nl = 0;
    return "tsar_slic: not built yet";
}
}  // namespace tsar
