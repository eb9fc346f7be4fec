/*
 * nis_ref_scaler.cpp -- NVScaler from the reference's own header, compiled on the host.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref).  build_ref.sh provides "NIS_Scaler_cpp.h": the file
 * /root/reference/src/nis/NIS_Scaler.h with ONE token changed (the HLSL float literal `255.0` of
 * NIS_SCALE_FLOAT spelled `255.0f`, because an unsuffixed literal is a double in C++ and a float in HLSL).
 * Configuration as in src/nis/NIS_Upscale.hlsl:22-26 except NIS_THREAD_GROUP_SIZE = 1: every loop in the
 * header strides by blockDim, so a single "thread" walks the whole 32x24 block and the barrier is a no-op.
 */
#include "hlsl_shim.h"

namespace nis_scaler_ref {
#include "hlsl_intrinsics.inc"
NIS_CB_FIELDS(NIS_CB_DECL)
static thread_local SamplerState samplerLinearClamp;
static thread_local Texture2D in_texture, coef_scaler, coef_usm;
static thread_local RWTexture2D out_texture;

#define NIS_SCALER 1
#define NIS_HDR_MODE 0
#define NIS_BLOCK_WIDTH 32
#define NIS_BLOCK_HEIGHT 24
#define NIS_THREAD_GROUP_SIZE 1
#include "NIS_Scaler_cpp.h"

static void bind(const ovo_image *src, const ovo_image *dst, const ovo_nis_config *cfg, const float *cs, const float *cu) {
  NIS_CB_FIELDS(NIS_CB_LOAD)
  in_texture.img = src; out_texture.img = dst; coef_scaler.table = cs; coef_usm.table = cu;
}
} // namespace nis_scaler_ref

extern "C" const float *ref_coef_scale(void);
extern "C" const float *ref_coef_usm(void);

#define OVO_ENTRY(n) ref_nis_scaler_##n
#define OVO_NIS_IS_SHARPEN 0
#define OVO_NIS_BIND(src, dst, cfg) nis_scaler_ref::bind(src, dst, cfg, ref_coef_scale(), ref_coef_usm())
#define OVO_NIS_BLOCK(src, dst, cfg, bx, by) nis_scaler_ref::NVScaler(uint2(bx, by), 0)
#include "../nis_entry.inc"

extern "C" int ref_nis_scaler(const ovo_image *src, const ovo_image *dst, const ovo_nis_config *c, int nthreads) {
  return ref_nis_scaler_run(src, dst, c, nthreads);
}
