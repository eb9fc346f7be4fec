/*
 * nis_ref_sharpen.cpp -- NVSharpen from the reference's own header, compiled on the host.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref); see nis_ref_scaler.cpp.  Configuration as in
 * src/nis/NIS_Sharpen.hlsl:22-26 (NIS_SCALER 0, 32x32 blocks) with NIS_THREAD_GROUP_SIZE = 1.
 */
#include "hlsl_shim.h"

namespace nis_sharpen_ref {
#include "hlsl_intrinsics.inc"
NIS_CB_FIELDS(NIS_CB_DECL)
static thread_local SamplerState samplerLinearClamp;
static thread_local Texture2D in_texture;
static thread_local RWTexture2D out_texture;

#define NIS_SCALER 0
#define NIS_HDR_MODE 0
#define NIS_BLOCK_WIDTH 32
#define NIS_BLOCK_HEIGHT 32
#define NIS_THREAD_GROUP_SIZE 1
#include "NIS_Scaler_cpp.h"

static void bind(const ovo_image *src, const ovo_image *dst, const ovo_nis_config *cfg) {
  NIS_CB_FIELDS(NIS_CB_LOAD)
  in_texture.img = src; out_texture.img = dst;
}
} // namespace nis_sharpen_ref

#define OVO_ENTRY(n) ref_nis_sharpen_##n
#define OVO_NIS_IS_SHARPEN 1
#define OVO_NIS_BIND(src, dst, cfg) nis_sharpen_ref::bind(src, dst, cfg)
#define OVO_NIS_BLOCK(src, dst, cfg, bx, by) nis_sharpen_ref::NVSharpen(uint2(bx, by), 0)
#include "../nis_entry.inc"

extern "C" int ref_nis_sharpen(const ovo_image *src, const ovo_image *dst, const ovo_nis_config *c, int nthreads) {
  return ref_nis_sharpen_run(src, dst, c, nthreads);
}
