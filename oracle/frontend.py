"""TEST INFRASTRUCTURE ONLY (see oracle/README or ovr_oracle.h): numpy restatement of the callers either side of the
hot path.  Only tests/ may import this.  Citations are /root/reference paths.

The reference has no tests or fixtures for any of this, and the D3D11 pieces (ResolveSubresource,
SaveDDSTextureToFile) are runtime/library behaviour rather than code in the tree, so these restate the public
definitions: ResolveSubresource = per-channel mean of the samples; DDS = the public container layout with the
choices ScreenGrab11.cpp makes for each DXGI format.  parity unpinned beyond that.
"""
from __future__ import annotations

import struct
import time

import numpy as np

FMT_RGBA8, FMT_BGRA8, FMT_RGBA16F, FMT_RGBA32F, FMT_RGB10A2 = 0, 1, 2, 3, 4


def recommended_render_size(fsr_enabled: bool, render_scale: float, width: int, height: int):
    """VrHooks.cpp:44-47: `*pnWidth *= renderScale` on uint32_t = (uint32)((float)w * scale), only when enabled
    and renderScale < 1."""
    s = np.float32(render_scale)
    if fsr_enabled and s < 1:
        return int(np.float32(width) * s), int(np.float32(height) * s)
    return width, height


def mip_lod_bias(input_width: int, output_width: int) -> np.float32:
    """PostProcessor.cpp:538: -log2(outputWidth / (float)inputWidth), binary32."""
    return np.float32(-np.log2(np.float32(output_width) / np.float32(input_width), dtype=np.float32))


def sampler_lod_bias(sampler_bias: float, max_anisotropy: int, bias: float) -> np.float32:
    """VrHooks.cpp:123-128."""
    b = np.float32(sampler_bias)
    return np.float32(b + np.float32(bias)) if (b == 0 and max_anisotropy > 1) else b


def capture_filename(use_nis: bool, sharpness: float, radius: float, unix_time: int) -> str:
    """PostProcessor.cpp:641-652 (roundf = half away from zero)."""
    def roundf(x):
        x = np.float32(x)
        return int(np.floor(x + np.float32(0.5))) if x >= 0 else -int(np.floor(-x + np.float32(0.5)))
    stamp = time.strftime("%Y%m%d_%H%M%S", time.localtime(unix_time))
    return "capture_%s_%s_s%d_r%d.dds" % (stamp, "nis" if use_nis else "fsr", roundf(np.float32(sharpness) * np.float32(100)),
                                          roundf(np.float32(radius) * np.float32(100)))


def _decode(a: np.ndarray, fmt: int) -> np.ndarray:
    """(H, N, 4) storage -> float32 channel values in storage order (no BGRA swap: the mean is channel-wise)."""
    if fmt in (FMT_RGBA8, FMT_BGRA8):
        return a.astype(np.float32) / np.float32(255.0)
    if fmt == FMT_RGB10A2:
        w = np.ascontiguousarray(a).view("<u4").reshape(a.shape[0], a.shape[1])
        c = np.stack([w & 1023, (w >> 10) & 1023, (w >> 20) & 1023], -1).astype(np.float32) / np.float32(1023.0)
        return np.concatenate([c, ((w >> 30).astype(np.float32) / np.float32(3.0))[..., None]], -1)
    return a.astype(np.float32)


def _encode(f: np.ndarray, fmt: int) -> np.ndarray:
    if fmt in (FMT_RGBA8, FMT_BGRA8):
        s = np.where(np.isnan(f), np.float32(0), np.clip(f, 0, 1)).astype(np.float32)
        return (s * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    if fmt == FMT_RGB10A2:
        s = np.where(np.isnan(f), np.float32(0), np.clip(f, 0, 1)).astype(np.float32)
        c = (s[..., :3] * np.float32(1023.0) + np.float32(0.5)).astype(np.uint32)
        a = (s[..., 3] * np.float32(3.0) + np.float32(0.5)).astype(np.uint32)
        w = c[..., 0] | (c[..., 1] << 10) | (c[..., 2] << 20) | (a << 30)
        return np.ascontiguousarray(w.astype("<u4")).view(np.uint8).reshape(f.shape[0], f.shape[1], 4)
    return f.astype(np.float16 if fmt == FMT_RGBA16F else np.float32)


def resolve_msaa(samples_img: np.ndarray, sample_count: int, fmt: int) -> np.ndarray:
    """GetInputView -> ResolveSubresource (PostProcessor.cpp:219-226): the standard resolve of a colour target =
    per-channel arithmetic mean of the texel's samples.  Stated here as: decode every sample exactly, add them in
    sample order in binary32, divide by the count (IEEE), store like any UNORM / float write."""
    h, n = samples_img.shape[:2]
    w = n // sample_count
    f = _decode(samples_img, fmt).reshape(h, w, sample_count, 4)
    acc = f[:, :, 0, :].copy()
    for i in range(1, sample_count):
        acc = (acc + f[:, :, i, :]).astype(np.float32)
    return _encode((acc / np.float32(sample_count)).astype(np.float32), fmt)


def dds_bytes(img: np.ndarray, fmt: int) -> bytes:
    """The file SaveDDSTextureToFile writes for a 2-D, single-mip texture (ScreenGrab11.cpp:815-935): magic, the
    124-byte header with flags TEXTURE|MIPMAP|PITCH, pitch = tight row bytes, mipMapCount 1, caps TEXTURE; pixel
    format = legacy masks for RGBA8 (A8B8G8R8, :168-169,837) / BGRA8 (A8R8G8B8, :162-163,857), D3DFMT FourCC for
    RGBA16F (113, :865) / RGBA32F (116, :864), 'DX10' + extension header for RGB10A2 (:204-208,875-883)."""
    h, w = img.shape[:2]
    bpp = {FMT_RGBA16F: 8, FMT_RGBA32F: 16}.get(fmt, 4)
    ext = b""
    if fmt == FMT_RGBA8:
        pf = struct.pack("<8I", 32, 0x41, 0, 32, 0x000000FF, 0x0000FF00, 0x00FF0000, 0xFF000000)
    elif fmt == FMT_BGRA8:
        pf = struct.pack("<8I", 32, 0x41, 0, 32, 0x00FF0000, 0x0000FF00, 0x000000FF, 0xFF000000)
    elif fmt in (FMT_RGBA16F, FMT_RGBA32F):
        pf = struct.pack("<8I", 32, 0x4, 113 if fmt == FMT_RGBA16F else 116, 0, 0, 0, 0, 0)
    else:
        pf = struct.pack("<8I", 32, 0x4, int.from_bytes(b"DX10", "little"), 0, 0, 0, 0, 0)
        ext = struct.pack("<5I", 24, 3, 0, 1, 0)  # DXGI_FORMAT_R10G10B10A2_UNORM, TEXTURE2D, misc 0, array 1
    head = struct.pack("<7I", 124, 0x1007 | 0x20000 | 0x8, h, w, w * bpp, 0, 1) + b"\0" * 44 + pf + \
        struct.pack("<5I", 0x1000, 0, 0, 0, 0)
    assert len(head) == 124
    return b"DDS " + head + ext + np.ascontiguousarray(img).tobytes()
