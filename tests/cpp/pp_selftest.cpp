// pp_selftest.cpp -- drives the C++ drop-in (vr::PostProcessor + Config singleton) exactly like the Submit detour
// does (/root/reference/src/postprocess/VrHooks.cpp:50-62): build a Texture_t, call Apply, read the swapped handle.
// usage: pp_selftest <in.rgba> <w> <h> <renderScale> <sharpness> <radius> <useNis> <out_left.rgba> <out_right.rgba> [captureDir]
// Built and run by tests/test_gpu_cpp_dropin.py (nvcc, links the in-tree libovrfsr.so).
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <vector>

#include <cuda_runtime.h>

#include "postprocessor.h"

static int fail(const char *what) { std::fprintf(stderr, "pp_selftest: %s\n", what); return 1; }

int main(int argc, char **argv) {
  if (argc != 10 && argc != 11) return fail("bad arguments");
  const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
  std::vector<unsigned char> host((size_t)w * h * 4);
  FILE *f = std::fopen(argv[1], "rb");
  if (!f || std::fread(host.data(), 1, host.size(), f) != host.size()) return fail("cannot read input");
  std::fclose(f);

  std::ostringstream log;
  SetLogStream(&log);
  Config &cfg = Config::Instance(); // what Config::Load() fills from openvr_mod.cfg in the mod
  cfg.renderScale = (float)std::atof(argv[4]);
  cfg.sharpness = (float)std::atof(argv[5]);
  cfg.radius = (float)std::atof(argv[6]);
  cfg.useNis = std::atoi(argv[7]) != 0;
  cfg.strictMath = true;

  ovrfsr_image eye[2];
  for (int e = 0; e < 2; ++e) {
    if (ovrfsr_image_alloc(&eye[e], w, h, OVRFSR_FORMAT_RGBA8) != OVRFSR_OK) return fail("image alloc (no GPU?)");
    // right eye = left eye shifted by 16 px, like the python tests
    std::vector<unsigned char> img(host);
    if (e == 1)
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
          for (int c = 0; c < 4; ++c) img[((size_t)y * w + x) * 4 + c] = host[((size_t)y * w + (x + w - 16 % w) % w) * 4 + c];
    if (cudaMemcpy2D(eye[e].data, eye[e].pitch, img.data(), (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice) != cudaSuccess)
      return fail("upload");
  }

  vr::PostProcessor postProcessor; // VrHooks.cpp:19
  vr::VRTextureBounds_t bounds{0.f, 0.f, 1.f, 1.f};

  // 1. fsrEnabled == false: Apply must leave the texture alone (PostProcessor.cpp:134)
  cfg.fsrEnabled = false;
  vr::Texture_t tex{&eye[0], vr::TextureType_OvrFsrCuda, vr::ColorSpace_Gamma};
  postProcessor.Apply(vr::Eye_Left, &tex, &bounds, vr::Submit_Default);
  if (tex.handle != &eye[0]) return fail("disabled post-processor modified the texture");
  // 2. a texture type it does not own is ignored (:124)
  cfg.fsrEnabled = true;
  vr::Texture_t gl{&eye[0], vr::TextureType_OpenGL, vr::ColorSpace_Auto};
  postProcessor.Apply(vr::Eye_Left, &gl, &bounds, vr::Submit_Default);
  if (gl.handle != &eye[0]) return fail("foreign texture type was processed");

  // 3. the hot path, both eyes, twice (second frame reuses the cached resources), then Reset and once more
  for (int frame = 0; frame < 3; ++frame) {
    if (frame == 2) postProcessor.Reset(); // ShutdownHooks / hotkeys
    for (int e = 0; e < 2; ++e) {
      vr::Texture_t t{&eye[e], vr::TextureType_OvrFsrCuda, vr::ColorSpace_Gamma};
      void *origHandle = t.handle;                                    // VrHooks.cpp:51
      postProcessor.Apply(e ? vr::Eye_Right : vr::Eye_Left, &t, &bounds, vr::Submit_Default);
      if (t.handle == origHandle) { std::fputs(log.str().c_str(), stderr); return fail("Apply did not swap the handle"); }
      if (t.eColorSpace != vr::ColorSpace_Gamma) return fail("colour space tag not preserved for an sRGB submit");
      const ovrfsr_image *out = static_cast<const ovrfsr_image *>(t.handle);
      if (frame == 2) {
        std::vector<unsigned char> res((size_t)out->width * out->height * 4);
        if (cudaMemcpy2D(res.data(), (size_t)out->width * 4, out->data, out->pitch, (size_t)out->width * 4, out->height,
                         cudaMemcpyDeviceToHost) != cudaSuccess)
          return fail("download");
        FILE *o = std::fopen(argv[8 + e], "wb");
        if (!o || std::fwrite(res.data(), 1, res.size(), o) != res.size()) return fail("cannot write output");
        std::fclose(o);
        std::printf("eye %d: %ux%u\n", e, out->width, out->height);
      }
      t.handle = origHandle;                                          // VrHooks.cpp:60
    }
  }
  // 4. F7: the next left-eye frame is written as a DDS file the library can read back (PostProcessor.cpp:634-657)
  if (argc > 10) {
    postProcessor.TakeCapture(argv[10]);
    vr::Texture_t t{&eye[0], vr::TextureType_OvrFsrCuda, vr::ColorSpace_Gamma};
    postProcessor.Apply(vr::Eye_Left, &t, &bounds, vr::Submit_Default);
    const ovrfsr_image *out = static_cast<const ovrfsr_image *>(t.handle);
    std::printf("capture requested: %ux%u\n", out->width, out->height);
  }
  // 5. the colour-space tag handed on to the real Submit (PostProcessor.cpp:162,504): Gamma stays Gamma; Auto becomes
  // Gamma exactly for the formats IsConsideredSrgbByOpenVR names (:76-92), carried as tag bits of the image format
  struct { int32_t fmt; vr::EColorSpace in, want; } cases[] = {
      {OVRFSR_FORMAT_RGBA8, vr::ColorSpace_Auto, vr::ColorSpace_Auto},
      {OVRFSR_FORMAT_RGBA8 | OVRFSR_FORMAT_SRGB_BIT, vr::ColorSpace_Auto, vr::ColorSpace_Gamma},
      {OVRFSR_FORMAT_RGBA8 | OVRFSR_FORMAT_TYPELESS_BIT, vr::ColorSpace_Auto, vr::ColorSpace_Gamma},
      {OVRFSR_FORMAT_RGBA8 | OVRFSR_FORMAT_SRGB_BIT, vr::ColorSpace_Linear, vr::ColorSpace_Auto},
      {OVRFSR_FORMAT_RGBA8, vr::ColorSpace_Gamma, vr::ColorSpace_Gamma},
  };
  for (const auto &c : cases) {
    postProcessor.Reset(); // inputIsSrgb is decided when the resources are (re)created, like the reference
    ovrfsr_image tagged = eye[0];
    tagged.format = c.fmt;
    vr::Texture_t t{&tagged, vr::TextureType_OvrFsrCuda, c.in};
    postProcessor.Apply(vr::Eye_Left, &t, &bounds, vr::Submit_Default);
    if (t.handle == &tagged) return fail("tagged format was not processed");
    if (t.eColorSpace != c.want) return fail("colour-space tag differs from PostProcessor.cpp:162,504");
  }
  for (int e = 0; e < 2; ++e) ovrfsr_image_free(&eye[e]);
  return 0;
}
