/*
 * cas_ref.cpp -- compiles the REFERENCE's own CAS kernel source lines on the host.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/build_ref.sh into oracle/_ref/ when /root/reference is present; nothing
 * from the reference is copied into the repository: build_ref.sh extracts
 *     src/cas/ffx_cas.h:409-893   (CasFilter, the non-packed float path; guarded by A_GPU in the reference)
 *     src/cas/ffx_a.h:1455-1457   (APrxLoSqrtF1, APrxLoRcpF1, APrxMedRcpF1)
 * into a temporary directory (rewriting HLSL "out T x" to "T& x").  The lines are compiled twice, as the two entry
 * shaders do: with CAS_BETTER_DIAGONALS for cas.sharpen.hlsl and without for cas.upscale.hlsl.
 */
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../ovr_glue.h"

namespace ref {

#include "amd_shim.h"
namespace { // internal linkage: fsr_ref.cpp holds the same-named functions of src/fsr/ffx_a.h
#include "cas_prx.inc" // the reference's APrxLoSqrtF1 / APrxLoRcpF1 / APrxMedRcpF1
}

static thread_local const ovo_image *g_tex = nullptr;
// cas.compute.h:14-19
static inline AF3 CasLoad(ASU2 p) { float t[4]; ovo_load(g_tex, p.x, p.y, t); return AF3(t[0], t[1], t[2]); }
static inline void CasInput(AF1 &, AF1 &, AF1 &) {}

namespace sharpen { // cas.sharpen.hlsl:1-3
#define CAS_BETTER_DIAGONALS 1
#include "cas_lines.inc"
#undef CAS_BETTER_DIAGONALS
} // namespace sharpen
namespace upscale { // cas.upscale.hlsl:1-2
#include "cas_lines.inc"
} // namespace upscale

static inline AU4 mk4(const uint32_t v[4]) { AU4 r; r.x = v[0]; r.y = v[1]; r.z = v[2]; r.w = v[3]; return r; }

static inline void cas_px(const ovo_image *src, const ovo_cas_constants *c, int sharpen_only, int x, int y, float o[4]) {
  g_tex = src;
  const AU2 gxy((uint32_t)x, (uint32_t)y);
  if (sharpen_only) sharpen::CasFilter(o[0], o[1], o[2], gxy, mk4(c->const0), mk4(c->const1), true);
  else upscale::CasFilter(o[0], o[1], o[2], gxy, mk4(c->const0), mk4(c->const1), false);
}

#define OVO_ENTRY(n) ref_cas_##n
#define OVO_CAS_PIXEL(src, c, so, x, y, o) cas_px(src, c, so, x, y, o)
#include "../cas_entry.inc"

} // namespace ref

extern "C" int ref_cas(const ovo_image *src, const ovo_image *dst, const ovo_cas_constants *c, int sharpen_only, int nthreads) {
  return ref::ref_cas_run_cas(src, dst, c, sharpen_only, nthreads);
}
