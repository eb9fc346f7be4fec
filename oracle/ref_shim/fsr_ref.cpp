/*
 * fsr_ref.cpp -- compiles the REFERENCE's own FSR1 kernel source lines on the host.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/build_ref.sh into oracle/_ref/ when
 * /root/reference is present; nothing from the reference is copied into the
 * repository: build_ref.sh extracts
 *     src/fsr/ffx_fsr1.h:239-437  (FsrEasuTapF, FsrEasuSetF, FsrEasuF)
 *     src/fsr/ffx_fsr1.h:684-769  (FsrRcasF)
 *     src/fsr/ffx_a.h:1843-1845   (APrxLoRcpF1, APrxMedRcpF1, APrxLoRsqF1)
 * into a temporary directory (rewriting HLSL "inout T x"/"out T x" to "T& x"),
 * and this file #includes them between the HLSL-type shim below and the entry
 * glue of ../fsr_entry.inc.  Those header sections are guarded by A_GPU in the
 * reference (ffx_fsr1.h:232,679) because they need HLSL vector types; the shim
 * supplies exactly those types, nothing of the algorithm.
 */
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../ovr_glue.h"

namespace ref {

#include "amd_shim.h"

#include "ffx_prx.inc" // the reference's APrxLoRcpF1 / APrxMedRcpF1 / APrxLoRsqF1

// ---- the texture callbacks the entry shaders provide (fsr_easu.hlsl:21-23, fsr_rcas.hlsl:18-19) ----
static thread_local const ovo_image *g_tex = nullptr;

// Gather4 at normalised p: footprint (i0,j0)..(i0+1,j0+1) around texel-space (p*size-0.5),
// returned as x=(i0,j0+1) y=(i0+1,j0+1) z=(i0+1,j0) w=(i0,j0), clamp-to-edge.
static inline AF4 gather(AF2 p, int ch) {
  const int i0 = (int)floorf(p.x * (float)g_tex->width - 0.5f);
  const int j0 = (int)floorf(p.y * (float)g_tex->height - 0.5f);
  float t[4];
  AF4 r;
  ovo_texel_clamp(g_tex, i0, j0 + 1, t); r.x = t[ch];
  ovo_texel_clamp(g_tex, i0 + 1, j0 + 1, t); r.y = t[ch];
  ovo_texel_clamp(g_tex, i0 + 1, j0, t); r.z = t[ch];
  ovo_texel_clamp(g_tex, i0, j0, t); r.w = t[ch];
  return r;
}
static inline AF4 FsrEasuRF(AF2 p) { return gather(p, 0); }
static inline AF4 FsrEasuGF(AF2 p) { return gather(p, 1); }
static inline AF4 FsrEasuBF(AF2 p) { return gather(p, 2); }
static inline AF4 FsrRcasLoadF(ASU2 p) { float t[4]; ovo_load(g_tex, p.x, p.y, t); return AF4(t[0], t[1], t[2], t[3]); }
static inline void FsrRcasInputF(AF1 &, AF1 &, AF1 &) {}

#define FSR_RCAS_LIMIT (0.25 - (1.0 / 16.0)) // ffx_fsr1.h:654 (a #define outside the extracted range)

#include "easu_lines.inc" // ffx_fsr1.h:239-437 verbatim modulo inout/out -> &
#include "rcas_lines.inc" // ffx_fsr1.h:684-769 verbatim modulo inout/out -> &

static inline AU4 mk4(const uint32_t v[4]) { AU4 r; r.x = v[0]; r.y = v[1]; r.z = v[2]; r.w = v[3]; return r; }

static inline void easu_px(const ovo_image *src, const ovo_upscale_constants *c, int x, int y, float o[4]) {
  g_tex = src;
  AF3 pix;
  FsrEasuF(pix, AU2((uint32_t)x, (uint32_t)y), mk4(c->const0), mk4(c->const1), mk4(c->const2), mk4(c->const3));
  o[0] = pix.x; o[1] = pix.y; o[2] = pix.z;
}
static inline void rcas_px(const ovo_image *src, const ovo_sharpen_constants *c, int x, int y, float o[4]) {
  g_tex = src;
  FsrRcasF(o[0], o[1], o[2], AU2((uint32_t)x, (uint32_t)y), mk4(c->const0));
}

#define OVO_ENTRY(n) ref_fsr_##n
#define OVO_EASU_PIXEL(src, c, x, y, o) easu_px(src, c, x, y, o)
#define OVO_RCAS_PIXEL(src, c, x, y, o) rcas_px(src, c, x, y, o)
#include "../fsr_entry.inc"

} // namespace ref

extern "C" {
int ref_fsr_easu(const ovo_image *src, const ovo_image *dst, const ovo_upscale_constants *c, int nthreads) {
  return ref::ref_fsr_run_easu(src, dst, c, nthreads);
}
int ref_fsr_rcas(const ovo_image *src, const ovo_image *dst, const ovo_sharpen_constants *c, int nthreads) {
  return ref::ref_fsr_run_rcas(src, dst, c, nthreads);
}
}
