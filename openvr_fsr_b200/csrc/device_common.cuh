// device_common.cuh -- pixel-format decode/encode and small math helpers shared by the
// FSR and NIS kernels (sm_100a).  Product code: never includes anything from oracle/.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ovrfsr.h"

// The math mode of this translation unit: the same sources are compiled twice,
//   kernels_fast.cu   : -fmad=true,  kStrict=false  (FMA contraction + regrouped taps)
//   kernels_strict.cu : -fmad=false, kStrict=true   (reference operation order, bit-exact)
// Every device symbol lives in a mode-specific inline namespace so that the two builds of the same
// template (e.g. easu_kernel<0,0>) get DIFFERENT mangled names; without it the linker folds the two
// weak instantiations into one and both launchers start the same kernel.
#ifndef OVRFSR_STRICT
#error "compile through kernels_fast.cu / kernels_strict.cu"
#endif
#if OVRFSR_STRICT
#define OVRFSR_MODE_NS strict_math
#else
#define OVRFSR_MODE_NS fast_math
#endif

namespace ovrfsr {
inline namespace OVRFSR_MODE_NS {

constexpr bool kStrict = (OVRFSR_STRICT != 0);

struct ImageRO { const uint8_t *ptr; uint32_t pitch; int w, h; };
struct ImageRW { uint8_t *ptr; uint32_t pitch; int w, h; };

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }

// ffx_a.h:1843-1845 -- integer bit-trick approximations (bit-exact on any IEEE machine)
__device__ __forceinline__ float prx_lo_rcp(float a) { return u2f(0x7ef07ebbu - f2u(a)); }
__device__ __forceinline__ float prx_lo_rsq(float a) { return u2f(0x5f347d74u - (f2u(a) >> 1)); }
__device__ __forceinline__ float prx_med_rcp(float a) {
  float b = u2f(0x7ef19fffu - f2u(a));
  return b * (-b * a + 2.0f);
}

// rcp(): IEEE 1/x in strict mode (what the checker computes), MUFU.RCP (<=1 ulp, D3D's own
// tolerance for rcp) in fast mode.
__device__ __forceinline__ float rcp_mode(float a) {
  if constexpr (kStrict) {
    return __frcp_rn(a);
  } else {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); // one MUFU.RCP; operands here are never subnormal
    return r;
  }
}

// Strict-math reciprocal for RCAS when the source is UNORM8: the operands are 4*mx and 4*mn-4 with mx, mn in
// {k/255}, i.e. one of 512 known values (or +0), all normal and far from overflow.  For those, MUFU.RCP plus one FMA
// Newton step IS the correctly rounded 1/x -- checked exhaustively on the device against rcp.rn by
// ovrfsr_selftest_rcp (tests/test_gpu_fsr_parity.py) -- without __frcp_rn's range checks and slow-path branch.
__device__ __forceinline__ float rcp_rn_unorm8_operand(float a) {
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(a));
  const float e = __fmaf_rn(-a, r0, 1.0f);
  const float r = __fmaf_rn(r0, e, r0);
  return a == 0.0f ? r0 : r; // 1/+0 = +inf (r would be NaN); -0 cannot occur here
}
template <bool UNORM8_OPERANDS>
__device__ __forceinline__ float rcp_mode_rcas(float a) {
  if constexpr (kStrict && UNORM8_OPERANDS) return rcp_rn_unorm8_operand(a);
  else return rcp_mode(a);
}

// byte k of a packed texel -> exact float(v): PRMT builds 0x4B0000vv (= 2^23 + v), one FADD removes 2^23.
// (A single I2F.U8 with a byte selector does the same on the conversion pipe; measured neutral: EASU 104.8 vs 104.9 us,
// RCAS 37.2 vs 37.2 us, so the two-instruction form that stays off that quarter-rate pipe is kept.)
template <int K>
__device__ __forceinline__ float byte_to_float(uint32_t p) {
  return u2f(__byte_perm(p, 0x4B000000u, 0x7540 | K)) - 8388608.0f;
}
// UNORM8 decode = correctly rounded v/255 (D3D11 UNORM->FLOAT) without a divide.  1/255 sits almost exactly half way
// between two floats; with chi the one below it and clo = float(1/255 - chi), fma(v, clo, v * chi) is the correctly
// rounded quotient for all 256 inputs (exhaustive exact-arithmetic check in tests/test_host_logic.py) -- one operation
// fewer than the residual-corrected multiply unorm10() uses (no such pair of constants exists for 1/1023).
__device__ __forceinline__ float unorm8(float v) {
  const float chi = __uint_as_float(0x3b808080u), clo = __uint_as_float(0x2f808081u);
  return __fmaf_rn(v, clo, __fmul_rn(v, chi));
}

// UNORM10 / UNORM2 decode of DXGI_FORMAT_R10G10B10A2_UNORM: the same residual-corrected multiply, exact (= v/1023, v/3
// correctly rounded) for all 1024 / 4 codes (tests/test_host_logic.py): q = v*r, one FMA residual step.
__device__ __forceinline__ float unorm10(uint32_t v) {
  const float r = 1.0f / 1023.0f, f = (float)v;
  const float q = __fmul_rn(f, r);
  return __fmaf_rn(__fmaf_rn(-1023.0f, q, f), r, q);
}
__device__ __forceinline__ float unorm2(uint32_t v) {
  const float r = 1.0f / 3.0f, f = (float)v;
  const float q = __fmul_rn(f, r);
  return __fmaf_rn(__fmaf_rn(-3.0f, q, f), r, q);
}
__device__ __forceinline__ float4 decode_rgb10a2(uint32_t p) {
  return make_float4(unorm10(p & 1023u), unorm10((p >> 10) & 1023u), unorm10((p >> 20) & 1023u), unorm2(p >> 30));
}
// formats stored as one 32-bit word per texel (the TMA tile loaders handle exactly these)
__host__ __device__ constexpr bool packed32(int fmt) {
  return fmt == OVRFSR_FORMAT_RGBA8 || fmt == OVRFSR_FORMAT_BGRA8 || fmt == OVRFSR_FORMAT_RGB10A2 || fmt == OVRFSR_FORMAT_BGRX8;
}

// Fast-math decode of byte K: PRMT builds the float 2^23+v, one FFMA computes (2^23+v)*r - 2^23*r = fl(v*r) with a
// single rounding (2^23*r is exact).  Differs from the correctly rounded v/255 by at most 1 ulp (126 of 256 codes);
// strict math always uses unorm8().
template <int K>
__device__ __forceinline__ float byte_to_unorm_mode(uint32_t p) {
  if constexpr (kStrict) {
    return unorm8(byte_to_float<K>(p));
  } else {
    const float r = 1.0f / 255.0f;
    return __fmaf_rn(u2f(__byte_perm(p, 0x4B000000u, 0x7540 | K)), r, -8388608.0f * r);
  }
}

// texel fetch -> float4 rgba.  No bounds logic here.
template <int FMT>
__device__ __forceinline__ float4 fetch_texel(const uint8_t *__restrict__ row, int x) {
  if constexpr (FMT == OVRFSR_FORMAT_RGBA32F) {
    return __ldg(reinterpret_cast<const float4 *>(row) + x);
  } else if constexpr (FMT == OVRFSR_FORMAT_RGBA16F) {
    const uint2 p = __ldg(reinterpret_cast<const uint2 *>(row) + x);
    const __half2 lo = *reinterpret_cast<const __half2 *>(&p.x), hi = *reinterpret_cast<const __half2 *>(&p.y);
    const float2 a = __half22float2(lo), b = __half22float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
  } else if constexpr (FMT == OVRFSR_FORMAT_RGB10A2) {
    return decode_rgb10a2(__ldg(reinterpret_cast<const uint32_t *>(row) + x));
  } else {
    const uint32_t p = __ldg(reinterpret_cast<const uint32_t *>(row) + x);
    const float c0 = unorm8(byte_to_float<0>(p)), c1 = unorm8(byte_to_float<1>(p));
    const float c2 = unorm8(byte_to_float<2>(p)), c3 = unorm8(byte_to_float<3>(p));
    if constexpr (FMT == OVRFSR_FORMAT_BGRA8) return make_float4(c2, c1, c0, c3);
    return make_float4(c0, c1, c2, c3);
  }
}

// FLOAT -> UNORM8: (uint)(saturate(v)*255 + 0.5); saturate(NaN) = 0 like D3D.
__device__ __forceinline__ uint32_t to_unorm8(float v) {
  float s = __saturatef(v);
  float t;
  if constexpr (kStrict) t = __fadd_rn(__fmul_rn(s, 255.0f), 0.5f);
  else t = __fmaf_rn(s, 255.0f, 0.5f);
  return (uint32_t)t; // F2I.TRUNC; t is in [0.5, 255.5]
}
// FLOAT -> UNORM with MAXV = 2^n - 1 (1023, 3): (uint)(saturate(v)*MAXV + 0.5)
template <int MAXV>
__device__ __forceinline__ uint32_t to_unorm(float v) {
  float s = __saturatef(v);
  float t;
  if constexpr (kStrict) t = __fadd_rn(__fmul_rn(s, (float)MAXV), 0.5f);
  else t = __fmaf_rn(s, (float)MAXV, 0.5f);
  return (uint32_t)t;
}

template <int FMT>
__device__ __forceinline__ void store_texel(uint8_t *__restrict__ row, int x, float r, float g, float b, float a) {
  if constexpr (FMT == OVRFSR_FORMAT_RGBA32F) {
    reinterpret_cast<float4 *>(row)[x] = make_float4(r, g, b, a);
  } else if constexpr (FMT == OVRFSR_FORMAT_RGBA16F) {
    const __half2 lo = __floats2half2_rn(r, g), hi = __floats2half2_rn(b, a);
    uint2 p;
    p.x = *reinterpret_cast<const uint32_t *>(&lo);
    p.y = *reinterpret_cast<const uint32_t *>(&hi);
    reinterpret_cast<uint2 *>(row)[x] = p;
  } else if constexpr (FMT == OVRFSR_FORMAT_RGB10A2) {
    reinterpret_cast<uint32_t *>(row)[x] = to_unorm<1023>(r) | (to_unorm<1023>(g) << 10) | (to_unorm<1023>(b) << 20) | (to_unorm<3>(a) << 30);
  } else {
    const uint32_t p = to_unorm8(r) | (to_unorm8(g) << 8) | (to_unorm8(b) << 16) | (to_unorm8(a) << 24);
    reinterpret_cast<uint32_t *>(row)[x] = p;
  }
}
// opaque RGBA8 pixel from three floats: floor(sat(v)*255 + 0.5) per channel (a 2^23 magic-add would round exact .5
// ties to even, and bilinear taps with 1/256 weights land on exact ties about 1% of the time), two byte permutes to pack.
__device__ __forceinline__ uint32_t pack_rgba8_opaque(float r, float g, float b) {
  return __byte_perm(__byte_perm(to_unorm8(r), to_unorm8(g), 0x0040), to_unorm8(b), 0x0410) | 0xff000000u;
}
// store (r,g,b,1): the common case of every filter on this path
template <int FMT>
__device__ __forceinline__ void store_opaque(uint8_t *__restrict__ row, int x, float r, float g, float b) {
  if constexpr (FMT == OVRFSR_FORMAT_RGBA8) reinterpret_cast<uint32_t *>(row)[x] = pack_rgba8_opaque(r, g, b);
  else if constexpr (FMT == OVRFSR_FORMAT_RGB10A2)
    reinterpret_cast<uint32_t *>(row)[x] = to_unorm<1023>(r) | (to_unorm<1023>(g) << 10) | (to_unorm<1023>(b) << 20) | 0xC0000000u;
  else store_texel<FMT>(row, x, r, g, b, 1.0f);
}

// Workgroup radius test (fsr_easu.hlsl:40-44, fsr_rcas.hlsl:31-35, NIS_Upscale.hlsl:98-101):
// wrapping u32 arithmetic against both eye centres at group granularity.
__device__ __forceinline__ bool group_inside(uint32_t gcx, uint32_t gcy, const uint32_t centre[4], uint32_t radiusSq) {
  const uint32_t d1x = centre[0] - gcx, d1y = centre[1] - gcy;
  const uint32_t d2x = centre[2] - gcx, d2y = centre[3] - gcy;
  return (d1x * d1x + d1y * d1y) <= radiusSq || (d2x * d2x + d2y * d2y) <= radiusSq;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// Packed FP32x2 (Blackwell FFMA2 / FMUL2 / FADD2): one issue slot for two FP32 lanes.  A scalar operand is
// broadcast for free (SASS `Rn.F32`), so `bc(w)` costs nothing.
typedef float2 f2;
__device__ __forceinline__ f2 bc(float x) { return make_float2(x, x); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false (it honours .rn only for scalar
// f32), and it also folds a literal "a*b + (-0)" / "a*1 + b" back into mul / add and then contracts those -- either
// would change results in the strict build (caught by the FP32-output parity tests).  The strict packed multiply / add
// are therefore FMAs against NEUTRAL ELEMENTS THE COMPILER CANNOT SEE: a __constant__ pair (-0, 1) that the host could
// overwrite, so it is loaded at run time.  fma(a, b, -0) = rn(a*b) and fma(a, 1, b) = rn(a+b) exactly (including
// the sign of zero), each a single rounding, and two real FMAs cannot be merged.
__constant__ float g_neutral[3] = {-0.0f, 1.0f, -1.0f};
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  if constexpr (kStrict) return __ffma2_rn(a, b, bc(g_neutral[0]));
  else return __fmul2_rn(a, b);
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  if constexpr (kStrict) return __ffma2_rn(a, bc(g_neutral[1]), b);
  else return __fadd2_rn(a, b);
}

// a - b, a*sa + b*sb and the HLSL lerp x + s*(y - x) on two lanes: in strict math every product and sum is its own
// rounding (explicit FMAs against the opaque neutral elements, see above); fast math contracts like the scalar code does
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  if constexpr (kStrict) return __ffma2_rn(b, bc(g_neutral[2]), a); // b * -1 is exact
  else return __ffma2_rn(b, bc(-1.0f), a);
}
__device__ __forceinline__ f2 madd2(f2 a, f2 sa, f2 b, f2 sb) {
  if constexpr (kStrict) return add2(mul2(a, sa), mul2(b, sb));
  else return fma2(b, sb, mul2(a, sa));
}
__device__ __forceinline__ f2 lerp2(f2 x, f2 y, f2 s) {
  if constexpr (kStrict) return add2(x, mul2(s, sub2(y, x)));
  else return fma2(s, sub2(y, x), x);
}

// Coordinate arithmetic is NEVER contracted, in either math mode: its result feeds floor()/int conversion, and a
// value that lands exactly on an integer (e.g. out x=140 of a 211->281 upscale maps to source 105.0) would pick a
// different texel footprint -- and a different de-ring clamp set -- if a*b+c were fused.  The reference leaves this
// to the shader compiler; the checker defines it as separate IEEE mul and add.
__device__ __forceinline__ float mul_add_unfused(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }

// D3D11 samplers convert the scaled texture coordinate to fixed point with 8 fractional bits
// (D3D11_SUBTEXEL_FRACTIONAL_BIT_COUNT): snap to 1/256 texel, round to nearest.  x256 is exact, so this is
// contraction-proof.
__device__ __forceinline__ float snap_subtexel(float s) { return floorf(fmaf(s, 256.0f, 0.5f)) * (1.0f / 256.0f); }

} // inline namespace OVRFSR_MODE_NS
} // namespace ovrfsr
