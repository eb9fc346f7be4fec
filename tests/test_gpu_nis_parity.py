"""GPU parity of the NIS alternate path (NVScaler / NVSharpen) against the oracle, through the C ABI.
strict -> bit-identical (checked at FP32 output, i.e. before any quantisation, and at RGBA8);
fast   -> <= 1 LSB per RGBA8 channel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [(33, 47, 0.75), (64, 48, 0.5), (100, 77, 0.77), (151, 93, 0.6), (40, 30, 1.3)]


def _imgs(w, h):
    from openvr_fsr_b200 import synth
    # black / grey: every pixel takes the edge-free shortcut (black also makes pixel_n * 255 a zero: the signed-zero case);
    # textured: warps with edge-free and edge pixels side by side
    black = np.zeros((h, w, 4), np.uint8)
    grey = np.full((h, w, 4), 77, np.uint8)
    black[..., 3] = grey[..., 3] = 255
    return {"uniform": synth.uniform_rgba8(w, h, 0), "natural": synth.natural_rgba8(w, h, 1),
            "textured": synth.textured_rgba8(w, h, 3, cell=16), "black": black, "grey": grey}


@pytest.mark.parametrize("iw,ih,scale", CASES)
@pytest.mark.parametrize("radius,sharp,debug", [(2.0, 0.9, False), (0.4, 0.3, True)])
def test_nvscaler_vs_oracle(cuda, iw, ih, scale, radius, sharp, debug):
    import torch
    import openvr_fsr_b200 as ovr
    from oracle import pyoracle as po
    ow, oh = po.output_size(iw, ih, scale)
    cfg, ok = po.nis_config(False, 0, True, iw, ih, ow, oh, radius=radius, sharpness=sharp, debug=debug)
    pcfg, pok = ovr.make_nis_config(ovr.Config(fsrEnabled=True, useNis=True, renderScale=scale, sharpness=sharp,
                                               radius=radius, debugMode=debug), False, 0, True, iw, ih, ow, oh)
    assert ok and pok and pcfg == bytes(cfg)
    for name, src in _imgs(iw, ih).items():
        t = torch.from_numpy(src).to(cuda)
        for odt, tdt in ((np.float32, torch.float32), (np.uint8, torch.uint8)):
            want = po.nis_scaler(src, ow, oh, cfg, out_dtype=odt)
            got = torch.zeros((oh, ow, 4), dtype=tdt, device=cuda)
            ovr.nis_scaler(t, got, pcfg, ovr.MATH_STRICT)
            torch.cuda.synchronize()
            assert np.array_equal(got.cpu().numpy().view(np.uint8), want.view(np.uint8)), (name, odt)
        fast = torch.zeros((oh, ow, 4), dtype=torch.uint8, device=cuda)
        ovr.nis_scaler(t, fast, pcfg, ovr.MATH_FAST)
        torch.cuda.synchronize()
        assert np.abs(fast.cpu().numpy().astype(np.int16) - want.astype(np.int16)).max() <= 1, name


@pytest.mark.parametrize("w,h", [(33, 47), (64, 64), (100, 77), (129, 31)])
@pytest.mark.parametrize("radius,sharp,debug", [(2.0, 0.9, False), (0.4, 0.3, True)])
def test_nvsharpen_vs_oracle(cuda, w, h, radius, sharp, debug):
    import torch
    import openvr_fsr_b200 as ovr
    from oracle import pyoracle as po
    cfg, ok = po.nis_config(True, 0, True, w, h, w, h, radius=radius, sharpness=sharp, debug=debug)
    pcfg, pok = ovr.make_nis_config(ovr.Config(fsrEnabled=True, useNis=True, renderScale=1.0, sharpness=sharp,
                                               radius=radius, debugMode=debug), True, 0, True, w, h, w, h)
    assert ok and pok and pcfg == bytes(cfg)
    for name, src in _imgs(w, h).items():
        src[..., 3] = np.random.default_rng(2).integers(0, 256, src.shape[:2], dtype=np.uint8)  # alpha passes through
        t = torch.from_numpy(src).to(cuda)
        for odt, tdt in ((np.float32, torch.float32), (np.uint8, torch.uint8)):
            want = po.nis_sharpen(src, cfg, out_dtype=odt)
            got = torch.zeros((h, w, 4), dtype=tdt, device=cuda)
            ovr.nis_sharpen(t, got, pcfg, ovr.MATH_STRICT)
            torch.cuda.synchronize()
            assert np.array_equal(got.cpu().numpy().view(np.uint8), want.view(np.uint8)), (name, odt)
        fast = torch.zeros((h, w, 4), dtype=torch.uint8, device=cuda)
        ovr.nis_sharpen(t, fast, pcfg, ovr.MATH_FAST)
        torch.cuda.synchronize()
        assert np.abs(fast.cpu().numpy().astype(np.int16) - want.astype(np.int16)).max() <= 1, name


def test_postprocessor_nis_paths(cuda):
    """useNis: renderScale != 1 -> NVScaler only; renderScale == 1 -> NVSharpen only (PostProcessor.cpp:586-594)."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih = 150, 110
    left, right = synth.stereo_pair("natural", iw, ih, 5)
    for scale in (0.75, 1.0):
        ow, oh = po.output_size(iw, ih, scale)
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, useNis=True, renderScale=scale, sharpness=0.9, radius=0.5,
                                          projCentre=(.47, .5, .53, .5)))
        for eye, img in ((0, left), (1, right)):
            got = pp.apply(eye, torch.from_numpy(img).to(cuda)).cpu().numpy()
            cfg, _ = po.nis_config(scale == 1.0, eye, True, iw, ih, ow, oh, proj=(.47, .5, .53, .5), radius=0.5, sharpness=0.9)
            want = po.nis_sharpen(img, cfg) if scale == 1.0 else po.nis_scaler(img, ow, oh, cfg)
            assert np.array_equal(got, want)
        pp.close()


def test_strict_nis_inrange_division_equals_div_rn(cuda):
    """Strict NIS divides with the range-check-free IEEE sequence when the source is UNORM; device check against
    div.rn over pseudo-random operands of the ranges that path can produce."""
    import ctypes as C
    from openvr_fsr_b200 import _lib as L
    bad, n = C.c_uint32(99), C.c_uint32(0)
    L.check(L.lib().ovrfsr_selftest_div(C.byref(bad), C.byref(n)))
    assert n.value == 2 * 64 * 148 * 4 * 256 and bad.value == 0
