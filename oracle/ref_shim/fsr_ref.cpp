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

// ---- HLSL scalar/vector types as the AMD headers name them (ffx_a.h A_HLSL section) ----
typedef float AF1;
typedef uint32_t AU1;
typedef bool AP1;

struct AF2 {
  float x, y;
  AF2() : x(0), y(0) {}
  AF2(double a, double b) : x((float)a), y((float)b) {}
  explicit AF2(const struct AU2 &u);
};
struct AF3pod { float r, g, b; };
struct AF3 {
  union { struct { float x, y, z; }; struct { float r, g, b; }; };
  AF3() : x(0), y(0), z(0) {}
  AF3(float a, float b_, float c) : x(a), y(b_), z(c) {}
  AF3(const AF3pod &p) : x(p.r), y(p.g), z(p.b) {}
};
struct AF4 {
  union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; AF3pod rgb; };
  AF4() : x(0), y(0), z(0), w(0) {}
  AF4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
};
struct AU2 { uint32_t x, y; AU2() : x(0), y(0) {} AU2(uint32_t a, uint32_t b) : x(a), y(b) {} };
struct AU2pod { uint32_t x, y; };
struct AU4 {
  union { struct { uint32_t x, y, z, w; }; struct { AU2pod xy, zw; }; };
};
struct ASU2 {
  int x, y;
  ASU2(int a, int b) : x(a), y(b) {}
  explicit ASU2(const AU2 &u) : x((int)u.x), y((int)u.y) {}
};
inline AF2::AF2(const AU2 &u) : x((float)u.x), y((float)u.y) {}

#define AF1_(a) ((ref::AF1)(a))
#define AU1_(a) ((ref::AU1)(a))
static inline AF2 AF2_(double a) { return AF2(a, a); }
static inline AF3 AF3_(float a) { return AF3(a, a, a); }
static inline AF4 AF4_(double a) { return AF4((float)a, (float)a, (float)a, (float)a); }
static inline AF1 AF1_AU1(AU1 u) { return ovo_u2f(u); }
static inline AU1 AU1_AF1(AF1 f) { return ovo_f2u(f); }
static inline AF2 AF2_AU2(const AU2pod &u) { AF2 r; r.x = ovo_u2f(u.x); r.y = ovo_u2f(u.y); return r; }

// ---- HLSL operators / intrinsics with D3D semantics (NaN-ignoring min/max, saturate(NaN)=0) ----
static inline AF2 operator+(AF2 a, AF2 b) { return AF2(a.x + b.x, a.y + b.y); }
static inline AF2 operator-(AF2 a, AF2 b) { return AF2(a.x - b.x, a.y - b.y); }
static inline AF2 operator*(AF2 a, AF2 b) { return AF2(a.x * b.x, a.y * b.y); }
static inline AF2 &operator-=(AF2 &a, AF2 b) { a.x -= b.x; a.y -= b.y; return a; }
static inline AF2 &operator*=(AF2 &a, AF2 b) { a.x *= b.x; a.y *= b.y; return a; }
static inline AF3 operator*(AF3 a, AF3 b) { return AF3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline AF3 operator*(AF3 a, float b) { return AF3(a.x * b, a.y * b, a.z * b); }
static inline AF3 &operator+=(AF3 &a, AF3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
static inline AF4 operator*(AF4 a, AF4 b) { return AF4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline AF4 operator+(AF4 a, AF4 b) { return AF4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline ASU2 operator+(ASU2 a, ASU2 b) { return ASU2(a.x + b.x, a.y + b.y); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float abs(float a) { return fabsf(a); }
static inline AF2 floor(AF2 a) { return AF2(floorf(a.x), floorf(a.y)); }
static inline AF3 min(AF3 a, AF3 b) { return AF3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
static inline AF3 max(AF3 a, AF3 b) { return AF3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
// ffx_a.h:1141,1143,1166,1168,1196,1206 (A_HLSL definitions; one-line wrappers around the intrinsics)
static inline AF1 AMax3F1(AF1 x, AF1 y, AF1 z) { return max(x, max(y, z)); }
static inline AF1 AMin3F1(AF1 x, AF1 y, AF1 z) { return min(x, min(y, z)); }
static inline AF3 AMax3F3(AF3 x, AF3 y, AF3 z) { return max(x, max(y, z)); }
static inline AF3 AMin3F3(AF3 x, AF3 y, AF3 z) { return min(x, min(y, z)); }
static inline AF1 ARcpF1(AF1 x) { return 1.0f / x; }                 // rcp(x)
static inline AF1 ASatF1(AF1 x) { return fminf(1.0f, fmaxf(0.0f, x)); } // saturate(x)

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
