// fast math mode: FMA contraction on, regrouped taps.  Build: nvcc -fmad=true
#define OVRFSR_STRICT 0
#include "kernels.inc"
