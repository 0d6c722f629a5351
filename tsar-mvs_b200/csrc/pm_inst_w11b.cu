// experimental launch shape for the 11x11 window: 256 threads, >= 3 CTAs/SM, unroll 2 (TSAR_B200_W11_VARIANT=b)
#define PM_FAST_UNROLL(n1) 2
#define PM_VARIANT pm_variant_w11b
#define PM_LABEL "w11b"
#define PM_NT 256
#define PM_MINB 3
#define PM_N1 6
#define PM_GEN false
#include "pm_inst.inc"
