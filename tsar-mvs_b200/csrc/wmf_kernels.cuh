// wmf_kernels.cuh -- weighted-median consistency filter (placeholder until the kernels land)
#pragma once
#include "glue_kernels.cuh"
namespace tsar {
static inline int wmf_launch(const GlueConst &, cudaTextureObject_t, float4 *, float *, float *, int, cudaStream_t) { return TSAR_ERR_STATE; }
static inline int wmf_final_launch(const GlueConst &, cudaTextureObject_t, float4 *, float *, float *, const float *, const float *, int, int, cudaStream_t) { return TSAR_ERR_STATE; }
}  // namespace tsar
