"""Source formats of PostProcessor.cpp:30-102 beyond the four RGBA layouts, on the CPU side: the checker's B8G8R8X8 and
R32G32B32_FLOAT fetches against their four-component equivalents, and the IsConsideredSrgbByOpenVR truth table."""
import numpy as np

from oracle import pyoracle as po
from openvr_fsr_b200 import synth


def test_oracle_bgrx8_reads_alpha_one():
    iw, ih, scale = 61, 45, 0.75
    ow, oh = po.output_size(iw, ih, scale)
    img = synth.uniform_rgba8(iw, ih, 3)
    img[..., 3] = np.random.default_rng(1).integers(0, 256, (ih, iw), dtype=np.uint8)  # junk in the X byte
    opaque = img.copy(); opaque[..., 3] = 255
    sc = po.sharpen_constants(0, True, iw, ih, radius=0.3, sharpness=0.8)          # mostly the outside-radius copy
    assert np.array_equal(po.rcas(img, sc, src_fmt=po.FMT_BGRX8), po.rcas(opaque, sc, src_fmt=po.FMT_BGRA8))
    assert not np.array_equal(po.rcas(img, sc, src_fmt=po.FMT_BGRA8), po.rcas(opaque, sc, src_fmt=po.FMT_BGRA8))
    cfg, _ = po.nis_config(False, 0, True, iw, ih, ow, oh, radius=2.0, sharpness=0.9)
    assert np.array_equal(po.nis_scaler(img, ow, oh, cfg, src_fmt=po.FMT_BGRX8), po.nis_scaler(opaque, ow, oh, cfg, src_fmt=po.FMT_BGRA8))
    scfg, _ = po.nis_config(True, 0, True, iw, ih, iw, ih, radius=2.0, sharpness=0.9)
    assert np.array_equal(po.nis_sharpen(img, scfg, src_fmt=po.FMT_BGRX8), po.nis_sharpen(opaque, scfg, src_fmt=po.FMT_BGRA8))


def test_oracle_rgb32f_equals_rgba32f_with_alpha_one():
    iw, ih, scale = 40, 33, 0.75
    ow, oh = po.output_size(iw, ih, scale)
    rgba = synth.natural_rgba16f(iw, ih, 2).astype(np.float32)
    rgba[..., 3] = 1.0
    rgb = np.ascontiguousarray(rgba[..., :3])
    uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=0.5)
    sc = po.sharpen_constants(0, True, ow, oh, radius=0.5, sharpness=0.9)
    a = po.rcas(po.easu(rgb, ow, oh, uc, src_fmt=po.FMT_RGB32F), sc)
    b = po.rcas(po.easu(rgba, ow, oh, uc), sc)
    assert np.array_equal(a, b)
    cfg, _ = po.nis_config(False, 0, True, iw, ih, ow, oh, radius=2.0, sharpness=0.9)
    assert np.array_equal(po.nis_scaler(rgb, ow, oh, cfg, src_fmt=po.FMT_RGB32F), po.nis_scaler(rgba, ow, oh, cfg))


def test_format_considered_srgb_truth_table():
    """IsConsideredSrgbByOpenVR, PostProcessor.cpp:76-92: the three 8-bit _SRGB formats, and the _TYPELESS variants of
    B8G8R8A8 / R8G8B8A8 / B8G8R8X8 / R10G10B10A2."""
    import openvr_fsr_b200 as ovr
    S, T = ovr.FORMAT_SRGB_BIT, ovr.FORMAT_TYPELESS_BIT
    eight = (ovr.FORMAT_RGBA8, ovr.FORMAT_BGRA8, ovr.FORMAT_BGRX8)
    for f in (ovr.FORMAT_RGBA8, ovr.FORMAT_BGRA8, ovr.FORMAT_RGBA16F, ovr.FORMAT_RGBA32F, ovr.FORMAT_RGB10A2, ovr.FORMAT_BGRX8,
              ovr.FORMAT_RGB32F):
        assert not ovr.format_considered_srgb(f)
        assert ovr.format_considered_srgb(f | S) == (f in eight)
        assert ovr.format_considered_srgb(f | T) == (f in eight or f == ovr.FORMAT_RGB10A2)
    assert not ovr.format_considered_srgb(ovr.FORMAT_AUTO)
