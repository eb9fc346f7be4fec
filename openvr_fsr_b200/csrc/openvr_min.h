// openvr_min.h -- the five OpenVR types vr::PostProcessor::Apply touches, re-declared as PODs so the
// drop-in builds without the 27k-line SDK headers (/root/reference/headers/openvr.h:149-153 EVREye,
// :155-166 ETextureType, :168-173 EColorSpace, :177-182 Texture_t, :609-613 VRTextureBounds_t,
// :639-668 EVRSubmitFlags).  If the real <openvr.h> was included first, its definitions are used and only
// the new texture-type tag below is added.
#pragma once
#include <stdint.h>

#ifndef _OPENVR_API
namespace vr {
enum EVREye { Eye_Left = 0, Eye_Right = 1 };
enum ETextureType {
  TextureType_Invalid = -1,
  TextureType_DirectX = 0,
  TextureType_OpenGL = 1,
  TextureType_Vulkan = 2,
  TextureType_IOSurface = 3,
  TextureType_DirectX12 = 4,
  TextureType_DXGISharedHandle = 5,
  TextureType_Metal = 6,
};
enum EColorSpace { ColorSpace_Auto = 0, ColorSpace_Gamma = 1, ColorSpace_Linear = 2 };
struct Texture_t {
  void *handle;
  ETextureType eType;
  EColorSpace eColorSpace;
};
struct VRTextureBounds_t { float uMin, vMin, uMax, vMax; };
enum EVRSubmitFlags {
  Submit_Default = 0x00,
  Submit_LensDistortionAlreadyApplied = 0x01,
  Submit_GlRenderBuffer = 0x02,
  Submit_Reserved = 0x04,
  Submit_TextureWithPose = 0x08,
  Submit_TextureWithDepth = 0x10,
};
} // namespace vr
#endif

namespace vr {
// Texture_t::eType value for "handle points at an ovrfsr_image (device-resident CUDA image)": the CUDA
// analogue of TextureType_DirectX + ID3D11Texture2D*.  Outside the SDK's enumerator range on purpose.
static const ETextureType TextureType_OvrFsrCuda = static_cast<ETextureType>(0x4f56);
} // namespace vr
