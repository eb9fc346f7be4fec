// dds_ref.cpp -- the reference's OWN DDS header construction (SaveDDSTextureToFile,
// src/postprocess/ScreenGrab11.cpp) as a host checker for openvr_fsr_b200/csrc/capture.cpp.
// TEST INFRASTRUCTURE ONLY.  build_ref.sh extracts, from the file where it lies,
//   dds_structs.inc = ScreenGrab11.cpp:72-208  (DDS_MAGIC, DDS_PIXELFORMAT / DDS_HEADER / DDS_HEADER_DXT10, flag
//                                               constants, the DDSPF_* pixel-format table) -- compiled verbatim
//   dds_setup.inc   = ScreenGrab11.cpp:819-887 (the header fill and the DXGI format switch) -- compiled verbatim
// and this shim supplies the few Windows / DXGI names those lines use.  Hand-written here: the uncompressed branch of
// :895-906 (flags |= DDS_HEADER_FLAGS_PITCH, pitchOrLinearSize = rowPitch) with the caller's tight row pitch, which is
// what GetSurfaceInfo (:456-560) returns for these formats ((width * bpp + 7) / 8).
#include <cstddef>
#include <cstdint>
#include <cstring>

// DXGI_FORMAT values (dxgiformat.h): the names the switch mentions; only the five the mod can capture are tested
enum DXGI_FORMAT : uint32_t {
  DXGI_FORMAT_R32G32B32A32_FLOAT = 2, DXGI_FORMAT_R16G16B16A16_FLOAT = 10, DXGI_FORMAT_R16G16B16A16_UNORM = 11,
  DXGI_FORMAT_R16G16B16A16_SNORM = 13, DXGI_FORMAT_R32G32_FLOAT = 16, DXGI_FORMAT_R10G10B10A2_UNORM = 24,
  DXGI_FORMAT_R8G8B8A8_UNORM = 28, DXGI_FORMAT_R8G8B8A8_SNORM = 31, DXGI_FORMAT_R16G16_FLOAT = 34, DXGI_FORMAT_R16G16_UNORM = 35,
  DXGI_FORMAT_R16G16_SNORM = 37, DXGI_FORMAT_R32_FLOAT = 41, DXGI_FORMAT_R8G8_UNORM = 49, DXGI_FORMAT_R8G8_SNORM = 51,
  DXGI_FORMAT_R16_FLOAT = 54, DXGI_FORMAT_R16_UNORM = 56, DXGI_FORMAT_R8_UNORM = 61, DXGI_FORMAT_A8_UNORM = 65,
  DXGI_FORMAT_R8G8_B8G8_UNORM = 68, DXGI_FORMAT_G8R8_G8B8_UNORM = 69, DXGI_FORMAT_BC1_UNORM = 71, DXGI_FORMAT_BC2_UNORM = 74,
  DXGI_FORMAT_BC3_UNORM = 77, DXGI_FORMAT_BC4_UNORM = 80, DXGI_FORMAT_BC4_SNORM = 81, DXGI_FORMAT_BC5_UNORM = 83,
  DXGI_FORMAT_BC5_SNORM = 84, DXGI_FORMAT_B5G6R5_UNORM = 85, DXGI_FORMAT_B5G5R5A1_UNORM = 86, DXGI_FORMAT_B8G8R8A8_UNORM = 87,
  DXGI_FORMAT_B8G8R8X8_UNORM = 88, DXGI_FORMAT_YUY2 = 107, DXGI_FORMAT_AI44 = 111, DXGI_FORMAT_IA44 = 112, DXGI_FORMAT_P8 = 113,
  DXGI_FORMAT_A8P8 = 114, DXGI_FORMAT_B4G4R4A4_UNORM = 115
};
#define D3D11_RESOURCE_DIMENSION_TEXTURE2D 3
#define MAKEFOURCC(ch0, ch1, ch2, ch3) \
  ((uint32_t)(uint8_t)(ch0) | ((uint32_t)(uint8_t)(ch1) << 8) | ((uint32_t)(uint8_t)(ch2) << 16) | ((uint32_t)(uint8_t)(ch3) << 24))
#define memcpy_s(dst, dstSize, src, n) memcpy((dst), (src), (n))
#define ERROR_NOT_SUPPORTED 50
#define HRESULT_FROM_WIN32(x) (-(int)(x))

namespace {
#include "dds_structs.inc" // ends inside the anonymous namespace; the pack pragma is re-balanced below
#pragma pack(pop)
} // namespace

extern "C" int ref_dds_header(uint8_t *out, uint32_t width, uint32_t height, uint32_t dxgiFormat, uint32_t rowPitch) {
  struct { uint32_t Width, Height; DXGI_FORMAT Format; } desc{width, height, static_cast<DXGI_FORMAT>(dxgiFormat)};
#include "dds_setup.inc"
  (void)extHeader;
  header->flags |= DDS_HEADER_FLAGS_PITCH;   /* ScreenGrab11.cpp:903 */
  header->pitchOrLinearSize = rowPitch;      /* :904 */
  memcpy(out, fileHeader, headerSize);
  return (int)headerSize;
}
