"""State machine of PostProcessor::Apply on the GPU (PostProcessor.cpp:123-164): lazy init, re-init on size change,
one shared texture for both eyes processed once per frame, array-slice right eye, debugMode timing, hotkey-style
config changes.  Results checked against the oracle (strict math)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_fsr(img, eye, one_eye, scale, radius, sharp, proj=(.5, .5, .5, .5), debug=False):
    from oracle import pyoracle as po
    ih, iw = img.shape[:2]
    ow, oh = po.output_size(iw, ih, scale)
    sc = po.sharpen_constants(eye, one_eye, ow, oh, proj=proj, radius=radius, sharpness=sharp, debug=debug)
    mid = img if scale == 1.0 else po.easu(img, ow, oh, po.upscale_constants(eye, one_eye, iw, ih, ow, oh, proj=proj, radius=radius))
    return po.rcas(mid, sc)


def test_shared_texture_is_processed_once_per_frame(cuda):
    """uMax-uMin == 0.5: both eyes live side by side in ONE texture; the first Submit of a frame processes it with
    two radius centres, the second Submit gets the same output without new work (PostProcessor.cpp:146,155-160,298-301)."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    tex_np = synth.natural_rgba8(320, 120, 21)
    tex = torch.from_numpy(tex_np).to(cuda)
    proj = (.45, .5, .55, .5)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.6, projCentre=proj))
    want = _oracle_fsr(tex_np, 0, False, 0.75, 0.6, 0.9, proj)
    for frame in range(2):
        n0 = ovr.kernel_launches()
        left = pp.apply(ovr.EYE_LEFT, tex, ovr.TextureBounds(0.0, 0.0, 0.5, 1.0))
        n1 = ovr.kernel_launches()
        right = pp.apply(ovr.EYE_RIGHT, tex, ovr.TextureBounds(0.5, 0.0, 1.0, 1.0))
        n2 = ovr.kernel_launches()
        assert n1 - n0 == 2 and n2 - n1 == 0  # EASU + RCAS once, nothing on the second Submit
        assert left.data_ptr() == right.data_ptr()
        assert np.array_equal(left.cpu().numpy(), want)
    pp.close()


def test_reinit_on_size_change_and_config_change(cuda):
    import dataclasses
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    cfg = ovr.Config(fsrEnabled=True, renderScale=0.77, sharpness=0.9, radius=0.5)
    pp = ovr.PostProcessor(cfg)
    for (w, h) in ((200, 120), (96, 140), (200, 120)):  # "Texture size changed, recreating resources" (:139-142)
        img = synth.natural_rgba8(w, h, w)
        got = pp.apply(0, torch.from_numpy(img).to(cuda)).cpu().numpy()
        assert np.array_equal(got, _oracle_fsr(img, 0, True, 0.77, 0.5, 0.9))
    # what the hotkeys do (:670-704): mutate the config, Reset
    img = synth.natural_rgba8(200, 120, 3)
    t = torch.from_numpy(img).to(cuda)
    for change in ({"sharpness": 0.4}, {"radius": 0.2}, {"debugMode": True}, {"renderScale": 1.0}):
        cfg = dataclasses.replace(cfg, **change)
        pp.set_config(cfg)
        got = pp.apply(0, t).cpu().numpy()
        assert np.array_equal(got, _oracle_fsr(img, 0, True, cfg.renderScale, cfg.radius, cfg.sharpness, debug=cfg.debugMode)), change
    assert np.array_equal(pp.upscale_constants(0)[16:], pp.sharpen_constants(0)[4:])
    pp.close()


def test_array_texture_right_eye_is_slice_1(cuda):
    """Texture arrays keep the right eye in slice 1 (PostProcessor.cpp:254-268)."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import _lib as L, synth
    iw, ih = 150, 90
    left, right = synth.stereo_pair("natural", iw, ih, 4)
    arr = torch.from_numpy(np.stack([left, right])).to(cuda)  # (2, H, W, 4): two slices
    lib = L.lib()
    cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=2.0).to_c()
    ctx = C.c_void_p()
    L.check(lib.ovrfsr_create(C.byref(ctx), C.byref(cfg)))
    src = L.Image(arr.data_ptr(), iw, ih, iw * 4, L.FORMAT_RGBA8, 2, iw * ih * 4)
    for eye, img in ((0, left), (1, right)):
        out = L.Image()
        L.check(lib.ovrfsr_apply(ctx, eye, C.byref(src), 1, C.byref(out), torch.cuda.current_stream().cuda_stream), "apply", ctx)
        torch.cuda.synchronize()
        got = ovr.api._wrap_device(out, cuda).cpu().numpy()
        assert np.array_equal(got, _oracle_fsr(img, eye, True, 0.75, 2.0, 0.9))
    lib.ovrfsr_destroy(ctx)


def test_debug_mode_collects_gpu_times(cuda):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    img = torch.from_numpy(synth.natural_rgba8(400, 300, 2)).to(cuda)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5, debugMode=True))
    for _ in range(12):
        pp.apply(0, img)
        torch.cuda.synchronize()
    ms, n = pp.gpu_time_ms()
    assert n >= 6 and 0.0 < ms < 50.0
    pp.close()


def test_apply_host_roundtrip(cuda):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    img = synth.natural_rgba8(301, 211, 8)
    ow, oh = po.output_size(301, 211, 0.75)
    src, dst = torch.from_numpy(img).pin_memory(), torch.empty((oh, ow, 4), dtype=torch.uint8).pin_memory()
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5))
    pp.apply_host(0, src, dst)
    torch.cuda.synchronize()
    assert np.array_equal(dst.numpy(), _oracle_fsr(img, 0, True, 0.75, 0.5, 0.9))
    pp.close()


def test_apply_host_pitched_host_rows(cuda):
    """Host images whose rows are padded take the 2-D copy path; the padding bytes must stay untouched."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    img = synth.natural_rgba8(301, 211, 9)
    ow, oh = po.output_size(301, 211, 0.75)
    sbuf = torch.zeros((211, 301 * 4 + 60), dtype=torch.uint8).pin_memory()
    dbuf = torch.full((oh, ow * 4 + 100), 0xAB, dtype=torch.uint8).pin_memory()
    src = torch.as_strided(sbuf, (211, 301, 4), (sbuf.stride(0), 4, 1))
    dst = torch.as_strided(dbuf, (oh, ow, 4), (dbuf.stride(0), 4, 1))
    src.copy_(torch.from_numpy(img))
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5))
    for eye in (0, 1, 0):  # also re-use of the staging buffers
        pp.apply_host(eye, src, dst)
        torch.cuda.synchronize()
        assert np.array_equal(dst.numpy(), _oracle_fsr(img, eye, True, 0.75, 0.5, 0.9))
        assert bool((dbuf[:, ow * 4:] == 0xAB).all())
    pp.close()


def test_apply_host_survives_a_size_change(cuda):
    """'Texture size changed, recreating resources' (PostProcessor.cpp:139-142) through the host entry: the reset must
    not free the eye that is being staged (round-1 advisor finding: the ctx disabled itself permanently)."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5))
    for (w, h, seed) in ((301, 211, 8), (160, 120, 9), (301, 211, 10)):
        img = synth.natural_rgba8(w, h, seed)
        ow, oh = po.output_size(w, h, 0.75)
        for eye in (0, 1):
            src, dst = torch.from_numpy(img).pin_memory(), torch.empty((oh, ow, 4), dtype=torch.uint8).pin_memory()
            pp.apply_host(eye, src, dst)
            torch.cuda.synchronize()
            assert np.array_equal(dst.numpy(), _oracle_fsr(img, eye, True, 0.75, 0.5, 0.9)), (w, h, eye)
    pp.close()
