/*
 * consts_ref.cpp -- the reference's own CPU-side constant setup, compiled as shipped.
 *
 * TEST INFRASTRUCTURE ONLY (built by oracle/build_ref.sh into oracle/_ref/).
 * Includes, from /root/reference/src, with A_CPU exactly as PostProcessor.cpp:7-11 does:
 *   fsr/ffx_a.h, fsr/ffx_fsr1.h  -> FsrEasuCon (ffx_fsr1.h:156), FsrRcasCon (:662)
 *   nis/NIS_Config.h             -> NVScalerUpdateConfig (:144), NVSharpenUpdateConfig (:244),
 *                                   coef_scale (:261), coef_usm (:328)
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#define A_CPU
#include "fsr/ffx_a.h"
#include "fsr/ffx_fsr1.h"
#include "nis/NIS_Config.h"

extern "C" {

void ref_FsrEasuCon(uint32_t con[16], float inVpW, float inVpH, float inW, float inH, float outW, float outH) {
  FsrEasuCon(con, con + 4, con + 8, con + 12, inVpW, inVpH, inW, inH, outW, outH);
}

void ref_FsrRcasCon(uint32_t con[4], float sharpnessStops) { FsrRcasCon(con, sharpnessStops); }

float ref_AClampF1(float x, float n, float m) { return AClampF1(x, n, m); }

/* fills the first 112 bytes of a 256-byte NISConfig; returns the bool the mod ignores */
int ref_NVScalerUpdateConfig(void *cfg256, float sharpness, uint32_t inW, uint32_t inH, uint32_t outW, uint32_t outH) {
  NISConfig c;
  memset(&c, 0, sizeof(c));
  /* argument pattern of PostProcessor.cpp:308 */
  bool ok = NVScalerUpdateConfig(c, sharpness, 0, 0, inW, inH, inW, inH, 0, 0, outW, outH, outW, outH);
  memcpy(cfg256, &c, sizeof(c));
  return ok ? 1 : 0;
}

int ref_NVSharpenUpdateConfig(void *cfg256, float sharpness, uint32_t inW, uint32_t inH) {
  NISConfig c;
  memset(&c, 0, sizeof(c));
  /* argument pattern of PostProcessor.cpp:433 */
  bool ok = NVSharpenUpdateConfig(c, sharpness, 0, 0, inW, inH, inW, inH, 0, 0);
  memcpy(cfg256, &c, sizeof(c));
  return ok ? 1 : 0;
}

uint32_t ref_sizeof_NISConfig(void) { return (uint32_t)sizeof(NISConfig); }
const float *ref_coef_scale(void) { return &coef_scale[0][0]; }
const float *ref_coef_usm(void) { return &coef_usm[0][0]; }
uint32_t ref_kPhaseCount(void) { return (uint32_t)kPhaseCount; }
uint32_t ref_kFilterSize(void) { return (uint32_t)kFilterSize; }
}
