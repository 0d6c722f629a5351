// 11x11 window (scripts: --blocksize=11): 36 samples.  128-thread CTAs, >= 4 CTAs/SM (128 registers, 36.9 KB of
// shared memory for the per-thread weight table), the whole window unrolled in the branch-free sampling loop
// (36 texture fetches in flight per thread).  Chosen by measurement on B200 (profiles/r01_variants_C2.json).
#define PM_FAST_UNROLL(n1) (n1)
#define PM_VARIANT pm_variant_w11
#define PM_LABEL "w11"
#define PM_NT 128
#define PM_MINB 4
#define PM_N1 6
#define PM_GEN false
#include "pm_inst.inc"
