// any window size / any cost combination: run-time loop bounds, reference-style insertion sort
#define PM_VARIANT pm_variant_generic
#define PM_LABEL "generic"
#define PM_NT 128
#define PM_MINB 1
#define PM_N1 0
#define PM_GEN true
#include "pm_inst.inc"
