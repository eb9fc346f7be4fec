/*
 * hlsl_shim.h -- the HLSL types, intrinsics and resource objects that NIS_Scaler.h expects from its host
 * shader (src/nis/NIS_Upscale.hlsl:22-74, src/nis/NIS_Sharpen.hlsl:22-72), so that the reference header
 * compiles VERBATIM as C++.  TEST INFRASTRUCTURE ONLY (part of oracle/_ref).  Nothing of the algorithm
 * lives here: vector structs, lerp/saturate/min/max with D3D semantics, a Texture2D that samples an
 * ovo_image through ../ovr_glue.h, a RWTexture2D proxy that stores through it.
 */
#pragma once
#include <cmath>
#include <cstdint>

#include "../ovr_glue.h"

typedef uint32_t uint;
struct float2 { float x, y; float2() : x(0), y(0) {} float2(float a, float b) : x(a), y(b) {} };
struct float3pod { float x, y, z; };
struct float3 {
  float x, y, z;
  float3() : x(0), y(0), z(0) {}
  float3(float a, float b, float c) : x(a), y(b), z(c) {}
  float3(const float3pod &p) : x(p.x), y(p.y), z(p.z) {}
};
struct float4 {
  union { struct { float x, y, z, w; }; float3pod xyz; float3pod rgb; };
  float4() : x(0), y(0), z(0), w(0) {}
  float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
struct int2 { int x, y; int2() : x(0), y(0) {} int2(int a, int b) : x(a), y(b) {} };
struct uint2 { uint x, y; uint2() : x(0), y(0) {} uint2(uint a, uint b) : x(a), y(b) {} };

static inline float4 operator*(const float4 &a, float s) { return float4(a.x * s, a.y * s, a.z * s, a.w * s); }

/* the intrinsics live in hlsl_intrinsics.inc, included INSIDE the namespace that holds the reference header so
 * that they hide (rather than overload) the C library's floor/ceil/abs. */

struct SamplerState {};
struct Texture2D {
  const ovo_image *img = nullptr; /* a colour texture ... */
  const float *table = nullptr;   /* ... or a 2x64 RGBA32F coefficient texture (PostProcessor.cpp:366-381) */
  float4 SampleLevel(SamplerState, float2 uv, int) const {
    float o[4];
    ovo_sample_linear(img, uv.x, uv.y, o);
    return float4(o[0], o[1], o[2], o[3]);
  }
  float4 operator[](int2 p) const { const float *t = table + p.y * 8 + p.x * 4; return float4(t[0], t[1], t[2], t[3]); }
};
struct RWTexture2D {
  const ovo_image *img = nullptr;
  struct Ref {
    const ovo_image *img; int x, y;
    void operator=(const float4 &v) const { const float c[4] = {v.x, v.y, v.z, v.w}; ovo_store(img, x, y, c); }
  };
  Ref operator[](uint2 p) const { return Ref{img, (int)p.x, (int)p.y}; }
};

#define groupshared static thread_local
#define NIS_UNROLL
static inline void GroupMemoryBarrierWithGroupSync() {}

/* the cbuffer of the host shader (NIS_Upscale.hlsl:28-66), one copy per worker thread */
#define NIS_CB_FIELDS(X)                                                                                      \
  X(float, kDetectRatio) X(float, kDetectThres) X(float, kMinContrastRatio) X(float, kRatioNorm)              \
  X(float, kContrastBoost) X(float, kEps) X(float, kSharpStartY) X(float, kSharpScaleY)                       \
  X(float, kSharpStrengthMin) X(float, kSharpStrengthScale) X(float, kSharpLimitMin) X(float, kSharpLimitScale) \
  X(float, kScaleX) X(float, kScaleY) X(float, kDstNormX) X(float, kDstNormY) X(float, kSrcNormX) X(float, kSrcNormY) \
  X(uint, kInputViewportOriginX) X(uint, kInputViewportOriginY) X(uint, kInputViewportWidth) X(uint, kInputViewportHeight) \
  X(uint, kOutputViewportOriginX) X(uint, kOutputViewportOriginY) X(uint, kOutputViewportWidth) X(uint, kOutputViewportHeight) \
  X(float, reserved0) X(float, reserved1)
#define NIS_CB_DECL(T, n) static thread_local T n;
#define NIS_CB_LOAD(T, n) n = cfg->n;
