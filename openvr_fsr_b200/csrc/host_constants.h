// host_constants.h -- CPU-side constant setup of the path (product code, header-only C++).
//
// Re-implements, for the CUDA backend, what PostProcessor::Prepare{Upscaling,Sharpening}Resources
// compute on the CPU before creating their immutable constant buffers
// (src/postprocess/PostProcessor.cpp:293-310,416-435,509-518 with src/fsr/ffx_fsr1.h:156-202,662-672
// and src/nis/NIS_Config.h:144-255; citations relative to /root/reference/).  The words produced
// here are compared bit-for-bit with the reference's own functions in tests/test_host_logic.py.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/ovrfsr.h"

namespace ovrfsr {
namespace host {

inline uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float from_bits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// PostProcessor.cpp:512-518: uint/float (or uint*float) evaluated in float, truncated to uint32
inline void output_size(uint32_t inW, uint32_t inH, float renderScale, uint32_t *outW, uint32_t *outH) {
  if (renderScale < 1.f) {
    *outW = static_cast<uint32_t>(static_cast<float>(inW) / renderScale);
    *outH = static_cast<uint32_t>(static_cast<float>(inH) / renderScale);
  } else {
    *outW = static_cast<uint32_t>(static_cast<float>(inW) * renderScale);
    *outH = static_cast<uint32_t>(static_cast<float>(inH) * renderScale);
  }
}

// FsrEasuCon (ffx_fsr1.h:156-202): each entry is x * (1/y) -- a multiply by the reciprocal.
inline void fsr_easu_con(uint32_t con[16], float vpW, float vpH, float inW, float inH, float outW, float outH) {
  const float rOutW = 1.0f / outW, rOutH = 1.0f / outH, rInW = 1.0f / inW, rInH = 1.0f / inH;
  const float v[16] = {vpW * rOutW,  vpH * rOutH,  0.5f * vpW * rOutW - 0.5f, 0.5f * vpH * rOutH - 0.5f,
                       rInW,         rInH,         1.0f * rInW,               -1.0f * rInH,
                       -1.0f * rInW, 2.0f * rInH,  1.0f * rInW,               2.0f * rInH,
                       0.0f * rInW,  4.0f * rInH,  0.0f,                      0.0f};
  for (int i = 0; i < 14; ++i) con[i] = bits(v[i]);
  con[14] = con[15] = 0;
}

// AU1_AH1_AF1 (ffx_a.h:482-550): truncating float->half that saturates to 0x7bff; expressed by its
// exponent rule rather than its 2x512-entry tables.
inline uint32_t half_bits_truncating(float f) {
  const uint32_t u = bits(f), e = (u >> 23) & 0xffu, sign = (u >> 16) & 0x8000u, man = u & 0x7fffffu;
  if (e < 103u) return sign + (man >> 24);
  if (e < 113u) return ((0x0400u >> (113u - e)) | sign) + (man >> (126u - e));
  if (e <= 142u) return (((e - 112u) << 10) | sign) + (man >> 13);
  return (0x7bffu | sign) + (man >> 24);
}

// FsrRcasCon (ffx_fsr1.h:662-672)
// CasSetup, src/cas/ffx_cas.h:375-397 with the A_CPU helpers of src/cas/ffx_a.h (ALerpF1 :302, ASatF1 :366, ARcpF1 :330)
inline void cas_setup(uint32_t c[8], float sharpness, float maxColorDelta, float inW, float inH, float outW, float outH) {
  c[0] = bits(inW * (1.0f / outW));
  c[1] = bits(inH * (1.0f / outH));
  c[2] = bits(0.5f * inW * (1.0f / outW) - 0.5f);
  c[3] = bits(0.5f * inH * (1.0f / outH) - 0.5f);
  const float lo = (0.0f > sharpness) ? 0.0f : sharpness, t = (1.0f < lo) ? 1.0f : lo;
  const float sharp = -(1.0f / (5.0f * t + (-8.0f * t + 8.0f)));
  c[4] = bits(sharp);
  c[5] = half_bits_truncating(sharp) + (half_bits_truncating(maxColorDelta) << 16);
  c[6] = bits(8.0f * inW * (1.0f / outW));
  c[7] = bits(maxColorDelta);
}

inline void fsr_rcas_con(uint32_t con[4], float stops) {
  const float s = exp2f(-stops);
  const uint32_t h = half_bits_truncating(s);
  con[0] = bits(s);
  con[1] = h + (h << 16);
  con[2] = con[3] = 0;
}

// imageCentre / radius words (PostProcessor.cpp:298-305 for the first buffer, :332-336 for the
// right-eye buffer of one-eye-per-texture submits)
inline void centre_radius(uint32_t centre[4], uint32_t radius[4], const ovrfsr_config &cfg, int eye, bool onlyOneEye,
                          uint32_t outW, uint32_t outH) {
  const float *proj = cfg.proj_centre;
  const float fw = static_cast<float>(outW), fh = static_cast<float>(outH);
  if (eye == 0 || !onlyOneEye) {
    const float half = static_cast<float>(outW / 2);
    centre[0] = static_cast<uint32_t>(onlyOneEye ? fw * proj[0] : half * proj[0]);
    centre[1] = static_cast<uint32_t>(fh * proj[1]);
    centre[2] = static_cast<uint32_t>(onlyOneEye ? fw * proj[0] : half * (1 + proj[2]));
    centre[3] = static_cast<uint32_t>(fh * (onlyOneEye ? proj[1] : proj[3]));
  } else {
    centre[0] = centre[2] = static_cast<uint32_t>(fw * proj[2]);
    centre[1] = centre[3] = static_cast<uint32_t>(fh * proj[3]);
  }
  radius[0] = static_cast<uint32_t>(0.5f * cfg.radius * fh);
  radius[1] = radius[0] * radius[0];
  radius[2] = outW;
  radius[3] = outH;
}

inline void make_upscale_constants(uint32_t c[24], const ovrfsr_config &cfg, int eye, bool onlyOneEye, uint32_t inW,
                                   uint32_t inH, uint32_t outW, uint32_t outH) {
  fsr_easu_con(c, (float)inW, (float)inH, (float)inW, (float)inH, (float)outW, (float)outH);
  centre_radius(c + 16, c + 20, cfg, eye, onlyOneEye, outW, outH);
}

inline void make_sharpen_constants(uint32_t c[12], const ovrfsr_config &cfg, int eye, bool onlyOneEye, uint32_t outW,
                                   uint32_t outH) {
  const float s = std::fmax(0.0f, std::fmin(cfg.sharpness, 1.0f)); // AClampF1, ffx_a.h:353
  fsr_rcas_con(c, 2.f - 2 * s);
  centre_radius(c + 4, c + 8, cfg, eye, onlyOneEye, outW, outH);
  c[3] = cfg.debug_mode ? 1u : 0u;
}

} // namespace host
} // namespace ovrfsr
