// experimental launch shape for the 11x11 window: 128 threads, >= 4 CTAs/SM, unroll (n1), warp = 8 columns (TSAR_B200_W11_VARIANT=f)
#define PM_FAST_UNROLL(n1) (n1)
#define PM_WARP_COLS 8
#define PM_VARIANT pm_variant_w11f
#define PM_LABEL "w11f"
#define PM_NT 128
#define PM_MINB 4
#define PM_N1 6
#define PM_GEN false
#include "pm_inst.inc"
