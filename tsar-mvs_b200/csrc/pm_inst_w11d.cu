// experimental launch shape for the 11x11 window: 256 threads, >= 2 CTAs/SM, half window per trip (TSAR_B200_W11_VARIANT=d)
#define PM_FAST_UNROLL(n1) (((n1) + 1) / 2)
#define PM_VARIANT pm_variant_w11d
#define PM_LABEL "w11d"
#define PM_NT 256
#define PM_MINB 2
#define PM_N1 6
#define PM_GEN false
#include "pm_inst.inc"
