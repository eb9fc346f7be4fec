"""BASELINE.json `configs` at FULL size, one eye each, CUDA path (through PostProcessor / the C ABI) vs the oracle.
strict math must be bit-identical; fast math within north_star's tolerance per pass (<=1 LSB RGBA8, <=1e-3 on FP16 out).
The oracle runs on all host threads (these sizes take a fraction of a second each that way)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1


def _lsb(a, b):
    return int(np.abs(a.astype(np.int16) - b.astype(np.int16)).max())


def test_c1_single_eye_960x1080_to_1920x2160_fsr(cuda):
    """configs[0]: single 960x1080 -> 1920x2160 RGBA8 eye, EASU+RCAS sharpness 0.9 (the CPU-runnable case)."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih, scale = 960, 1080, 0.5
    ow, oh = po.output_size(iw, ih, scale)
    assert (ow, oh) == (1920, 2160)
    src = synth.natural_rgba8(iw, ih, 1)
    uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=2.0)
    sc = po.sharpen_constants(0, True, ow, oh, radius=2.0, sharpness=0.9)
    easu = po.easu(src, ow, oh, uc, nthreads=NT)
    want = po.rcas(easu, sc, nthreads=NT)
    t = ovr.to_image(src, cuda)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=2.0, mathMode=ovr.MATH_STRICT))
    assert np.array_equal(pp.apply(0, t).cpu().numpy(), want)
    pp.close()
    fe = ovr.alloc_image(ow, oh, torch.uint8, cuda)
    fr = ovr.alloc_image(ow, oh, torch.uint8, cuda)
    ovr.fsr_easu(t, fe, uc.words(), ovr.MATH_FAST)
    ovr.fsr_rcas(ovr.to_image(easu, cuda), fr, sc.words(), ovr.MATH_FAST)
    torch.cuda.synchronize()
    assert _lsb(fe.cpu().numpy(), easu) <= 1 and _lsb(fr.cpu().numpy(), want) <= 1


def test_c3a_fp16_supersample_easu_rcas_radius_mask(cuda):
    """configs[2] as the reference can actually run it: FP16 2244x2492, renderScale 1.3 ->
    2917x3239 (the code's formula, not the README's 2915x3240), EASU+RCAS, radius 0.5, RGBA8 out."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih, scale = 2244, 2492, 1.3
    ow, oh = po.output_size(iw, ih, scale)
    assert (ow, oh) == (2917, 3239)
    src = synth.natural_rgba16f(iw, ih, 2)
    uc = po.upscale_constants(1, True, iw, ih, ow, oh, radius=0.5)
    sc = po.sharpen_constants(1, True, ow, oh, radius=0.5, sharpness=0.9)
    easu = po.easu(src, ow, oh, uc, nthreads=NT)
    want = po.rcas(easu, sc, nthreads=NT)
    t = ovr.to_image(src, cuda)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=0.5, mathMode=ovr.MATH_STRICT))
    assert np.array_equal(pp.apply(1, t).cpu().numpy(), want)
    pp.close()
    fe = ovr.alloc_image(ow, oh, torch.uint8, cuda)
    ovr.fsr_easu(t, fe, uc.words(), ovr.MATH_FAST)
    torch.cuda.synchronize()
    assert _lsb(fe.cpu().numpy(), easu) <= 1


def test_c3b_rcas_only_fp16_2915x3240_mask_on(cuda):
    """configs[2] read literally: RCAS only on a 2915x3240 FP16 frame, radius 0.5; RGBA8 out for the
    1-LSB bar and FP16 out for the 1e-3 bar."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    w, h = 2915, 3240
    src = synth.natural_rgba16f(w, h, 3, peak=1.0)
    sc = po.sharpen_constants(0, True, w, h, radius=0.5, sharpness=0.9)
    t = ovr.to_image(src, cuda)
    for odt, tdt in ((np.uint8, torch.uint8), (np.float16, torch.float16)):
        want = po.rcas(src, sc, out_dtype=odt, nthreads=NT)
        for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
            got = ovr.alloc_image(w, h, tdt, cuda)
            ovr.fsr_rcas(t, got, sc.words(), mode)
            torch.cuda.synchronize()
            g = got.cpu().numpy()
            if mode == ovr.MATH_STRICT:
                assert np.array_equal(g.view(np.uint8), want.view(np.uint8))
            elif odt == np.uint8:
                assert _lsb(g, want) <= 1
            else:
                assert float(np.abs(g.astype(np.float32) - want.astype(np.float32)).max()) <= 1e-3
    # and through PostProcessor: renderScale == 1 -> sharpen pass only (PostProcessor.cpp:586-594)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=1.0, sharpness=0.9, radius=0.5, mathMode=ovr.MATH_STRICT))
    assert np.array_equal(pp.apply(0, t).cpu().numpy(), po.rcas(src, sc, nthreads=NT))
    pp.close()


def test_c4_nis_1512x1680_to_2016x2240(cuda):
    """configs[3]: stereo 1512x1680 -> 2016x2240, useNIS: NVScaler only (no second pass, PostProcessor.cpp:591)."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih, scale = 1512, 1680, 0.75
    ow, oh = po.output_size(iw, ih, scale)
    assert (ow, oh) == (2016, 2240)
    left, right = synth.stereo_pair("natural", iw, ih, 4)
    for radius in (2.0, 0.5):
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, useNis=True, renderScale=scale, sharpness=0.9, radius=radius,
                                          mathMode=ovr.MATH_STRICT))
        for eye, img in ((0, left), (1, right)):
            cfg, ok = po.nis_config(False, eye, True, iw, ih, ow, oh, radius=radius, sharpness=0.9)
            assert ok
            want = po.nis_scaler(img, ow, oh, cfg, nthreads=NT)
            t = ovr.to_image(img, cuda)
            assert np.array_equal(pp.apply(eye, t).cpu().numpy(), want)
            fast = ovr.alloc_image(ow, oh, torch.uint8, cuda)
            ovr.nis_scaler(t, fast, bytes(cfg), ovr.MATH_FAST)
            torch.cuda.synchronize()
            assert _lsb(fast.cpu().numpy(), want) <= 1
        pp.close()


def test_c5_batched_frames_are_independent_of_sharding(cuda):
    """configs[4] in miniature: a batch of independent stereo frames gives identical bytes whether processed in one go
    or as the per-rank shards of a 4-way split (kernels are deterministic: no atomics, no reductions)."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import sharding, synth
    cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5)
    frames = [synth.stereo_pair("natural", 320, 200, 100 + i) for i in range(8)]

    def run(indices):
        pp = ovr.PostProcessor(cfg)
        out = {}
        for i in indices:
            out[i] = tuple(pp.apply(eye, ovr.to_image(frames[i][eye], cuda)).cpu().numpy().copy() for eye in (0, 1))
        pp.close()
        return out

    whole = run(range(8))
    for rank in range(4):
        part = run(sharding.frames_for_rank(8, rank, 4))
        for i, (l, r) in part.items():
            assert np.array_equal(l, whole[i][0]) and np.array_equal(r, whole[i][1])
