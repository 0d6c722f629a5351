// 19x19 window (AlgorithmParameters default, algorithmparameters.h:25-26): 100 samples, 128-thread CTAs;
// checkerboard kernel: w per thread + one reference tile per CTA (60 KB) -> 3 CTAs per SM instead of 2
// (64-thread CTAs x 7 per SM were measured too: 108 ms instead of 80 ms on C1 -- 128 registers are too few here.)
#define PM_VARIANT pm_variant_w19
#define PM_LABEL "w19"
#define PM_NT 128
#define PM_MINB 3
#define PM_TILE 1
#define PM_N1 10
#define PM_GEN false
#include "pm_inst.inc"
