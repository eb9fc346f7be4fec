"""CPU tests for the callers either side of the path: render-size / MIP-bias policy, the capture
file name, and the DDS container -- library (C ABI, no GPU needed) against the numpy restatement in oracle/frontend.py."""
import os
import time

import numpy as np
import pytest

import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
from oracle import frontend as fe


@pytest.mark.parametrize("scale", [0.5, 0.59, 0.67, 0.75, 0.77, 0.83, 0.99, 1.0, 1.3])
@pytest.mark.parametrize("enabled", [True, False])
def test_recommended_render_size(scale, enabled):
    cfg = ovr.Config(fsrEnabled=enabled, renderScale=scale)
    for w, h in ((2244, 2492), (1920, 2160), (2016, 2240), (1, 1), (4095, 4097), (3001, 3337)):
        assert ovr.recommended_render_size(cfg, w, h) == fe.recommended_render_size(enabled, scale, w, h)
    # the BASELINE shapes: a 2244x2492 recommendation at 0.75 renders at 1683x1869, which PrepareResources maps back
    if enabled and scale == 0.75:
        assert ovr.recommended_render_size(cfg, 2244, 2492) == (1683, 1869)
        assert ovr.output_size(1683, 1869, 0.75) == (2244, 2492)


def test_mip_lod_bias_and_sampler_rule():
    for iw, ow in ((1683, 2244), (960, 1920), (1512, 2016), (2244, 2917), (1000, 1000), (1323, 2244)):
        got, want = ovr.mip_lod_bias(iw, ow), fe.mip_lod_bias(iw, ow)
        assert abs(got - float(want)) <= 1.2e-7 * max(1.0, abs(float(want)))  # libm log2f: <= 1 ulp
    assert ovr.mip_lod_bias(960, 1920) == -1.0
    b = ovr.mip_lod_bias(1683, 2244)
    assert -0.42 < b < -0.41
    # VrHooks.cpp:123-128: only unbiased anisotropic samplers get the bias
    for sb, aniso in ((0.0, 16), (0.0, 2), (0.0, 1), (0.0, 0), (-0.5, 16), (1.0, 8)):
        assert ovr.sampler_lod_bias(sb, aniso, b) == float(fe.sampler_lod_bias(sb, aniso, b))
    assert ovr.sampler_lod_bias(0.0, 16, b) == b and ovr.sampler_lod_bias(0.0, 1, b) == 0.0 and ovr.sampler_lod_bias(-0.5, 16, b) == -0.5


def test_capture_filename():
    t = int(time.mktime((2021, 7, 14, 9, 5, 7, 0, 0, -1)))
    for nis, sharp, radius in ((False, 0.9, 0.5), (True, 0.75, 0.6), (False, 0.675, 0.125), (True, 1.0, 2.0), (False, 0.0, 0.005)):
        cfg = ovr.Config(fsrEnabled=True, useNis=nis, sharpness=sharp, radius=radius)
        assert ovr.capture_filename(cfg, t) == fe.capture_filename(nis, sharp, radius, t)
    assert ovr.capture_filename(ovr.Config(sharpness=0.9, radius=0.5), t) == "capture_20210714_090507_fsr_s90_r50.dds"


@pytest.mark.parametrize("fmt,make", [
    (fe.FMT_RGBA8, lambda: synth.natural_rgba8(37, 21, 2)),
    (fe.FMT_BGRA8, lambda: synth.uniform_rgba8(16, 9, 3)),
    (fe.FMT_RGBA16F, lambda: synth.natural_rgba16f(33, 17, 4)),
    (fe.FMT_RGBA32F, lambda: synth.natural_f32(20, 11, 5)),
    (fe.FMT_RGB10A2, lambda: synth.natural_rgb10a2(29, 13, 6)),
])
def test_dds_file_is_byte_identical_and_round_trips(tmp_path, fmt, make):
    img = make()
    if img.shape[-1] == 3:
        img = np.concatenate([img, np.ones_like(img[..., :1])], -1)
    path = tmp_path / "c.dds"
    ovr.save_dds(path, img, fmt)
    data = path.read_bytes()
    assert data == fe.dds_bytes(img, fmt)
    assert len(data) == 4 + 124 + (20 if fmt == fe.FMT_RGB10A2 else 0) + img.nbytes
    back, bfmt = ovr.load_dds(path)
    assert bfmt == fmt and back.dtype == img.dtype and np.array_equal(back.view(np.uint8), img.view(np.uint8))


def test_dds_pitched_source_and_errors(tmp_path):
    buf = np.zeros((9, 16 * 4 + 24), dtype=np.uint8)
    img = synth.uniform_rgba8(16, 9, 1)
    view = np.lib.stride_tricks.as_strided(buf, (9, 16, 4), (buf.strides[0], 4, 1))
    view[...] = img
    import ctypes as C
    from openvr_fsr_b200 import _lib as L
    im = L.Image(buf.ctypes.data, 16, 9, buf.strides[0], L.FORMAT_RGBA8, 1, 0, 0)
    p = tmp_path / "p.dds"
    assert L.lib().ovrfsr_dds_write(str(p).encode(), C.byref(im)) == L.OK
    assert p.read_bytes() == fe.dds_bytes(img, fe.FMT_RGBA8)  # rows are written tight
    # unreadable / foreign files fail cleanly
    (tmp_path / "junk.dds").write_bytes(b"DDS " + b"\0" * 40)
    with pytest.raises(ovr.OvrFsrError):
        ovr.load_dds(tmp_path / "junk.dds")
    with pytest.raises(ovr.OvrFsrError):
        ovr.load_dds(tmp_path / "missing.dds")
    dxt = bytearray(fe.dds_bytes(img, fe.FMT_RGBA8))
    dxt[76 + 4:76 + 12] = (0x4).to_bytes(4, "little") + b"DXT1"  # pixel format at file offset 76: flags = FOURCC, fourCC = DXT1
    (tmp_path / "dxt.dds").write_bytes(bytes(dxt))
    with pytest.raises(ovr.OvrFsrError) as e:
        ovr.load_dds(tmp_path / "dxt.dds")
    assert e.value.status == ovr.ERR_UNSUPPORTED
    with pytest.raises(ovr.OvrFsrError):
        ovr.save_dds(os.path.join(str(tmp_path), "no_such_dir", "x.dds"), img)


def test_resolve_restatement_properties():
    """The numpy resolve itself: identical samples resolve to themselves; a 2-sample 0/255 texel resolves to 128."""
    img = synth.uniform_rgba8(12, 7, 9)
    for s in (2, 4, 8):
        assert np.array_equal(fe.resolve_msaa(np.repeat(img, s, axis=1), s, fe.FMT_RGBA8), img)
    two = np.zeros((1, 2, 4), np.uint8)
    two[0, 1] = 255
    assert fe.resolve_msaa(two, 2, fe.FMT_RGBA8).tolist() == [[[128, 128, 128, 128]]]
    ten = synth.uniform_rgb10a2(10, 5, 2)
    assert np.array_equal(fe.resolve_msaa(np.repeat(ten, 4, axis=1), 4, fe.FMT_RGB10A2), ten)


def _dds_header_of_library(tmp_path, fmt, w, h, bpt):
    """Header bytes of the file the product's writer (capture.cpp, ovrfsr_dds_write) produces for a w x h image."""
    import ctypes as C
    from openvr_fsr_b200 import _lib as L
    buf = np.zeros((h, w * bpt), dtype=np.uint8)
    im = L.Image(buf.ctypes.data, w, h, w * bpt, fmt, 1, 0, 0)
    p = tmp_path / f"h_{fmt}_{w}x{h}.dds"
    assert L.lib().ovrfsr_dds_write(str(p).encode(), C.byref(im)) == L.OK
    data = p.read_bytes()
    return data[: len(data) - buf.nbytes]


def test_dds_header_equals_the_reference_writers(tmp_path):
    """Pins the capture container to the reference: tests/golden/dds_headers.json holds the headers produced by
    SaveDDSTextureToFile's own lines (ScreenGrab11.cpp:72-208,819-906 compiled into oracle/_ref, see
    tests/golden/make_golden_dds.py); the library's writer and the numpy restatement must reproduce them byte for byte."""
    import json
    from pathlib import Path
    from oracle import pyoracle as po
    cases = json.loads((Path(__file__).parent / "golden" / "dds_headers.json").read_text())
    assert len(cases) == 15
    for c in cases:
        want = bytes.fromhex(c["header_hex"])
        if c["width"] * c["height"] <= 1 << 20:  # the 2244x2492 header only from the cheap paths below
            got = _dds_header_of_library(tmp_path, c["format"], c["width"], c["height"], c["bytes_per_texel"])
            assert got == want, c
        dt = {fe.FMT_RGBA16F: np.float16, fe.FMT_RGBA32F: np.float32}.get(c["format"], np.uint8)
        if c["width"] * c["height"] <= 1 << 20:
            blob = fe.dds_bytes(np.zeros((c["height"], c["width"], 4), dt), c["format"])
            assert blob[: len(want)] == want, c
        if po.ref_available():  # the fixture itself against a fresh run of the reference lines
            assert po.ref_dds_header(c["width"], c["height"], c["format"], c["bytes_per_texel"]) == want
