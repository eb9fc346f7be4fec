// postprocessor.h -- C++ drop-in for vr::PostProcessor on the CUDA backend.
//
// Same public surface as /root/reference/src/postprocess/PostProcessor.h:12-13, so the Submit detours
// (src/postprocess/VrHooks.cpp:50-88) and ShutdownHooks (:155) compile against it unchanged:
//     postProcessor.Apply(eEye, pTexture, pBounds, nSubmitFlags);   postProcessor.Reset();
// Texture_t::handle carries an `ovrfsr_image*` (eType == TextureType_OvrFsrCuda) instead of an
// ID3D11Texture2D*.  All GPU work goes through the C ABI in include/ovrfsr.h.
#pragma once

#include <iosfwd>

#include "../../include/ovrfsr.h"
#include "config.h"
#include "openvr_min.h"

OVRFSR_API std::ostream &Log(); // Config.cpp:24-32 in the reference; here a settable stream (default std::clog)
OVRFSR_API void SetLogStream(std::ostream *os);

namespace vr {

class OVRFSR_API PostProcessor {
public:
  PostProcessor() = default;
  ~PostProcessor();
  PostProcessor(const PostProcessor &) = delete;
  PostProcessor &operator=(const PostProcessor &) = delete;

  void Apply(EVREye eEye, const Texture_t *pTexture, const VRTextureBounds_t *pBounds, EVRSubmitFlags nSubmitFlags);
  void Reset();

  // backend wiring the D3D11 version got from the texture itself (device/context, PostProcessor.cpp:500-501)
  void SetStream(void *cudaStream) { stream = cudaStream; }
  // mean GPU ms per frame in debugMode ("Average GPU processing time for upscale", PostProcessor.cpp:619-626)
  bool GetAverageGpuTimeMs(float *ms);
  // what the F7 hotkey does (`takeCapture = true`, PostProcessor.cpp:699-702): the next left-eye Apply writes its
  // output as capture_<time>_<fsr|nis>_s<..>_r<..>.dds into `directory` (the reference uses the DLL's directory)
  void TakeCapture(const char *directory);

private:
  bool enabled = true;
  bool initialized = false;
  bool inputIsSrgb = false;
  ovrfsr_ctx *ctx = nullptr;
  void *stream = nullptr;
  ovrfsr_image outputImage[2] = {};
  bool takeCapture = false;
  char captureDir[512] = {};
};

} // namespace vr
