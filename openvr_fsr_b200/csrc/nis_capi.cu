// nis_capi.cu -- NIS half of include/ovrfsr.h: NISConfig setup, filter banks, NVScaler / NVSharpen dispatch.
#include <cstring>
#include <mutex>

#include "kernels.h"
#include "nis_host.h"
#include "nis_coef_table.inc"

using namespace ovrfsr;

namespace {
float g_coefScale[64][8], g_coefUsm[64][8]; // reference layout: 64 phases x kFilterSize(8), 6 used
std::once_flag g_coefOnce;
void expand_coef() {
  for (int p = 0; p < 64; ++p)
    for (int t = 0; t < 8; ++t) {
      g_coefScale[p][t] = t < 6 ? (float)(kNisCoefScale1e4[p][t] / 10000.0) : 0.0f;
      g_coefUsm[p][t] = t < 6 ? (float)(kNisCoefUsm1e4[p][t] / 10000.0) : 0.0f;
    }
}
} // namespace

extern "C" {

int ovrfsr_make_nis_config(void *cfg256, const ovrfsr_config *cfg, int sharpen_only, int eye, int only_one_eye,
                           uint32_t in_w, uint32_t in_h, uint32_t out_w, uint32_t out_h) {
  if (!cfg256 || !cfg) return 0;
  host::NisConfig c;
  const bool ok = host::make_nis_config(c, *cfg, sharpen_only != 0, eye, only_one_eye != 0, in_w, in_h, out_w, out_h);
  std::memcpy(cfg256, &c, sizeof(c));
  return ok ? 1 : 0;
}

const float *ovrfsr_nis_coef_scale(void) { std::call_once(g_coefOnce, expand_coef); return &g_coefScale[0][0]; }
const float *ovrfsr_nis_coef_usm(void) { std::call_once(g_coefOnce, expand_coef); return &g_coefUsm[0][0]; }

} // extern "C"

// NVScaler / NVSharpen dispatches live in nis_kernels.cuh (added with the NIS kernels).
#ifndef OVRFSR_HAVE_NIS_KERNELS
extern "C" {
int ovrfsr_dispatch_nis_scaler(const ovrfsr_image *, const ovrfsr_image *, const void *, int, void *) {
  return OVRFSR_ERR_UNSUPPORTED;
}
int ovrfsr_dispatch_nis_sharpen(const ovrfsr_image *, const ovrfsr_image *, const void *, int, void *) {
  return OVRFSR_ERR_UNSUPPORTED;
}
}
#endif
