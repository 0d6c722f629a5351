// 19x19 window (AlgorithmParameters default, algorithmparameters.h:25-26): 100 samples, 128-thread CTAs
#define PM_VARIANT pm_variant_w19
#define PM_LABEL "w19"
#define PM_NT 128
#define PM_MINB 2
#define PM_N1 10
#define PM_GEN false
#include "pm_inst.inc"
