/*
 * cas_consts_ref.cpp -- the reference's CasSetup (src/cas/ffx_cas.h:375-397) compiled as shipped under A_CPU with its
 * own src/cas/ffx_a.h.  A separate translation unit because src/fsr/ffx_a.h (a newer revision of the same header)
 * is what consts_ref.cpp includes.  TEST INFRASTRUCTURE ONLY.
 */
#include <cstdint>
#include <cmath>
#define A_CPU 1
#include "cas/ffx_a.h"
#include "cas/ffx_cas.h"

extern "C" void ref_cas_setup(uint32_t const0[4], uint32_t const1[4], float sharpness, float maxColorDelta, float inW, float inH,
                              float outW, float outH) {
  CasSetup(const0, const1, sharpness, maxColorDelta, inW, inH, outW, outH);
}
