// 11x11 window (scripts: --blocksize=11): 36 samples, 256-thread CTAs, weight table 72 KB / CTA
#define PM_VARIANT pm_variant_w11
#define PM_LABEL "w11"
#define PM_NT 256
#define PM_MINB 2
#define PM_N1 6
#define PM_GEN false
#include "pm_inst.inc"
