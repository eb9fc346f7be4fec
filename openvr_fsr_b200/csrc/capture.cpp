// capture.cpp -- host-only callers either side of the path (/root/reference paths):
//   * the DDS file the F7 capture writes (PostProcessor.cpp:640-657 -> ScreenGrab11.cpp:815-935), writer AND
//     reader, so a capture from a real D3D11 box can be diffed against this library's output;
//   * the render-target-size and MIP-LOD-bias policy (VrHooks.cpp:37-48,123-128, PostProcessor.cpp:537-538).
// Nothing here touches the GPU.  The container layout follows the public DDS specification; only the choices
// SaveDDSTextureToFile makes for these formats (which legacy pixel format, which header flags) are mirrored.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/ovrfsr.h"

namespace {

constexpr uint32_t kMagic = 0x20534444u; // "DDS "
// header flags: CAPS | HEIGHT | WIDTH | PIXELFORMAT, MIPMAPCOUNT, PITCH (ScreenGrab11.cpp:96-98,827,882)
constexpr uint32_t kFlagsTexture = 0x00001007u, kFlagsMipmap = 0x00020000u, kFlagsPitch = 0x00000008u;
constexpr uint32_t kCapsTexture = 0x00001000u;
constexpr uint32_t kPfFourCC = 0x4u, kPfRGBA = 0x41u; // DDPF_FOURCC, DDPF_RGB | DDPF_ALPHAPIXELS
constexpr uint32_t kFourCCDX10 = 0x30315844u;          // 'D','X','1','0'
constexpr uint32_t kDxgiR10G10B10A2Unorm = 24u, kDxgiR8G8B8A8Unorm = 28u, kDxgiB8G8R8A8Unorm = 87u;
constexpr uint32_t kDxgiR16G16B16A16Float = 10u, kDxgiR32G32B32A32Float = 2u;
constexpr uint32_t kDimTexture2D = 3u;

struct PixelFormat { uint32_t size, flags, fourCC, bitCount, rMask, gMask, bMask, aMask; };
struct Header {
  uint32_t size, flags, height, width, pitchOrLinearSize, depth, mipMapCount, reserved1[11];
  PixelFormat pf;
  uint32_t caps, caps2, caps3, caps4, reserved2;
};
struct HeaderDX10 { uint32_t dxgiFormat, resourceDimension, miscFlag, arraySize, reserved; };
static_assert(sizeof(Header) == 124 && sizeof(PixelFormat) == 32 && sizeof(HeaderDX10) == 20, "DDS header layout");

uint32_t bpp_of(int fmt) { return fmt == OVRFSR_FORMAT_RGBA32F ? 16u : (fmt == OVRFSR_FORMAT_RGBA16F ? 8u : 4u); }

} // namespace

extern "C" {

int ovrfsr_dds_write(const char *path, const ovrfsr_image *im) {
  if (!path || !im || !im->data || im->width == 0 || im->height == 0) return OVRFSR_ERR_INVALID;
  Header h{};
  HeaderDX10 ext{};
  bool dx10 = false;
  h.size = sizeof(Header);
  h.flags = kFlagsTexture | kFlagsMipmap | kFlagsPitch;
  h.height = im->height;
  h.width = im->width;
  h.mipMapCount = 1;
  h.caps = kCapsTexture;
  h.pf.size = sizeof(PixelFormat);
  switch (im->format) {
    case OVRFSR_FORMAT_RGBA8: h.pf = {32, kPfRGBA, 0, 32, 0x000000ffu, 0x0000ff00u, 0x00ff0000u, 0xff000000u}; break; // A8B8G8R8
    case OVRFSR_FORMAT_BGRA8: h.pf = {32, kPfRGBA, 0, 32, 0x00ff0000u, 0x0000ff00u, 0x000000ffu, 0xff000000u}; break; // A8R8G8B8
    case OVRFSR_FORMAT_RGBA16F: h.pf.flags = kPfFourCC; h.pf.fourCC = 113; break; // D3DFMT_A16B16G16R16F
    case OVRFSR_FORMAT_RGBA32F: h.pf.flags = kPfFourCC; h.pf.fourCC = 116; break; // D3DFMT_A32B32G32R32F
    case OVRFSR_FORMAT_RGB10A2: // legacy 10:10:10:2 masks are ambiguous between writers: DX10 extension
      h.pf.flags = kPfFourCC; h.pf.fourCC = kFourCCDX10;
      dx10 = true;
      ext = {kDxgiR10G10B10A2Unorm, kDimTexture2D, 0, 1, 0};
      break;
    default: return OVRFSR_ERR_UNSUPPORTED;
  }
  const uint32_t rowBytes = im->width * bpp_of(im->format);
  if (im->pitch < rowBytes) return OVRFSR_ERR_INVALID;
  h.pitchOrLinearSize = rowBytes;
  FILE *f = std::fopen(path, "wb");
  if (!f) return OVRFSR_ERR_INVALID;
  bool ok = std::fwrite(&kMagic, 4, 1, f) == 1 && std::fwrite(&h, sizeof(h), 1, f) == 1;
  if (ok && dx10) ok = std::fwrite(&ext, sizeof(ext), 1, f) == 1;
  for (uint32_t y = 0; ok && y < im->height; ++y)
    ok = std::fwrite(static_cast<const uint8_t *>(im->data) + (size_t)y * im->pitch, 1, rowBytes, f) == rowBytes;
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) std::remove(path); // like ScreenGrab's auto_delete_file: no half-written captures
  return ok ? OVRFSR_OK : OVRFSR_ERR_INVALID;
}

int ovrfsr_dds_read(const char *path, ovrfsr_image *out) {
  if (!path || !out) return OVRFSR_ERR_INVALID;
  FILE *f = std::fopen(path, "rb");
  if (!f) return OVRFSR_ERR_INVALID;
  uint32_t magic = 0;
  Header h{};
  HeaderDX10 ext{};
  int rc = OVRFSR_OK, fmt = -1;
  if (std::fread(&magic, 4, 1, f) != 1 || magic != kMagic || std::fread(&h, sizeof(h), 1, f) != 1 || h.size != sizeof(Header) ||
      h.pf.size != sizeof(PixelFormat) || h.width == 0 || h.height == 0)
    rc = OVRFSR_ERR_INVALID;
  if (rc == OVRFSR_OK) {
    if (h.pf.flags & kPfFourCC) {
      if (h.pf.fourCC == 113) fmt = OVRFSR_FORMAT_RGBA16F;
      else if (h.pf.fourCC == 116) fmt = OVRFSR_FORMAT_RGBA32F;
      else if (h.pf.fourCC == kFourCCDX10) {
        if (std::fread(&ext, sizeof(ext), 1, f) != 1) rc = OVRFSR_ERR_INVALID;
        else if (ext.resourceDimension != kDimTexture2D || ext.arraySize > 1) rc = OVRFSR_ERR_UNSUPPORTED;
        else if (ext.dxgiFormat == kDxgiR10G10B10A2Unorm) fmt = OVRFSR_FORMAT_RGB10A2;
        else if (ext.dxgiFormat == kDxgiR8G8B8A8Unorm) fmt = OVRFSR_FORMAT_RGBA8;
        else if (ext.dxgiFormat == kDxgiB8G8R8A8Unorm) fmt = OVRFSR_FORMAT_BGRA8;
        else if (ext.dxgiFormat == kDxgiR16G16B16A16Float) fmt = OVRFSR_FORMAT_RGBA16F;
        else if (ext.dxgiFormat == kDxgiR32G32B32A32Float) fmt = OVRFSR_FORMAT_RGBA32F;
      }
    } else if (h.pf.bitCount == 32 && h.pf.gMask == 0x0000ff00u) {
      if (h.pf.rMask == 0x000000ffu && h.pf.bMask == 0x00ff0000u) fmt = OVRFSR_FORMAT_RGBA8;
      else if (h.pf.rMask == 0x00ff0000u && h.pf.bMask == 0x000000ffu) fmt = OVRFSR_FORMAT_BGRA8;
    }
    if (rc == OVRFSR_OK && fmt < 0) rc = OVRFSR_ERR_UNSUPPORTED;
  }
  if (rc == OVRFSR_OK) {
    const size_t rowBytes = (size_t)h.width * bpp_of(fmt), total = rowBytes * h.height;
    void *data = std::malloc(total);
    if (!data) rc = OVRFSR_ERR_NOMEM;
    else if (std::fread(data, 1, total, f) != total) { std::free(data); rc = OVRFSR_ERR_INVALID; }
    else *out = ovrfsr_image{data, h.width, h.height, (uint32_t)rowBytes, fmt, 1, 0, 0};
  }
  std::fclose(f);
  return rc;
}

void ovrfsr_host_free(void *p) { std::free(p); }

int ovrfsr_capture_filename(const ovrfsr_config *cfg, int64_t unix_time, char *buf, uint32_t n) {
  if (!cfg || !buf || n == 0) return OVRFSR_ERR_INVALID;
  char stamp[16];
  const std::time_t t = (std::time_t)unix_time;
  std::tm tmv{};
  localtime_r(&t, &tmv);
  std::strftime(stamp, sizeof(stamp), "%Y%m%d_%H%M%S", &tmv);
  const int w = std::snprintf(buf, n, "capture_%s_%s_s%d_r%d.dds", stamp, cfg->use_nis ? "nis" : "fsr",
                              (int)roundf(cfg->sharpness * 100), (int)roundf(cfg->radius * 100));
  return (w > 0 && (uint32_t)w < n) ? OVRFSR_OK : OVRFSR_ERR_INVALID;
}

void ovrfsr_recommended_render_size(const ovrfsr_config *cfg, uint32_t *width, uint32_t *height) {
  if (!cfg || !width || !height) return; // VrHooks.cpp:40-42
  if (cfg->fsr_enabled && cfg->render_scale < 1) {
    *width = (uint32_t)((float)*width * cfg->render_scale);   // `*pnWidth *= renderScale` on a uint32_t
    *height = (uint32_t)((float)*height * cfg->render_scale);
  }
}

float ovrfsr_mip_lod_bias(uint32_t input_width, uint32_t output_width) {
  return -log2f((float)output_width / (float)input_width);
}

float ovrfsr_sampler_lod_bias(float sampler_bias, uint32_t max_anisotropy, float mip_lod_bias) {
  return (sampler_bias == 0 && max_anisotropy > 1) ? sampler_bias + mip_lod_bias : sampler_bias;
}

} // extern "C"
