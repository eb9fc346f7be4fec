// strict math mode: reference operation order, no contraction.  Build: nvcc -fmad=false
#define OVRFSR_STRICT 1
#include "kernels.inc"
