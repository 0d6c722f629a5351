// experimental launch shape for the 11x11 window: 128 threads, >= 5 CTAs/SM, unroll (((n1) + 1) / 2) (TSAR_B200_W11_VARIANT=c)
#define PM_FAST_UNROLL(n1) (((n1) + 1) / 2)
#define PM_VARIANT pm_variant_w11c
#define PM_LABEL "w11c"
#define PM_NT 128
#define PM_MINB 5
#define PM_N1 6
#define PM_GEN false
#include "pm_inst.inc"
