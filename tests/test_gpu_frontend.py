"""GPU parity for the front-end and capture rows: MSAA resolve (GetInputView,
PostProcessor.cpp:219-226) alone and in front of the passes, and the F7 capture (PostProcessor.cpp:630-657)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _msaa(img, samples, seed, fmt):
    """a multisampled version of `img`: every sample = the texel plus a small per-sample perturbation"""
    rng = np.random.default_rng(seed)
    rep = np.repeat(img, samples, axis=1)
    if img.dtype == np.uint8 and fmt != 4:
        return np.clip(rep.astype(np.int16) + rng.integers(-20, 21, rep.shape), 0, 255).astype(np.uint8)
    if fmt == 4:
        from openvr_fsr_b200 import synth
        c = synth.unpack_rgb10a2(rep)
        c[..., :3] = np.clip(c[..., :3] + rng.integers(-60, 61, c[..., :3].shape), 0, 1023)
        c[..., 3] = rng.integers(0, 4, c[..., 3].shape)
        return synth.pack_rgb10a2(c[..., :3], c[..., 3])
    return (rep.astype(np.float32) * rng.uniform(0.9, 1.1, rep.shape).astype(np.float32)).astype(img.dtype)


@pytest.mark.parametrize("samples", [2, 4, 8])
@pytest.mark.parametrize("fmt", [0, 1, 2, 3, 4])
def test_resolve_vs_restatement(cuda, samples, fmt):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import frontend as fe
    w, h = 301, 57
    base = {0: synth.natural_rgba8(w, h, 1), 1: synth.uniform_rgba8(w, h, 2), 2: synth.natural_rgba16f(w, h, 3),
            3: synth.natural_rgba16f(w, h, 4).astype(np.float32), 4: synth.natural_rgb10a2(w, h, 5)}[fmt]
    ms = _msaa(base, samples, 10 + fmt, fmt)
    want = fe.resolve_msaa(ms, samples, fmt)
    src = torch.from_numpy(ms).to(cuda)
    dst = torch.zeros_like(torch.from_numpy(base)).to(cuda)
    ovr.resolve_msaa(src, dst, samples, fmt=fmt)
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy().view(np.uint8), want.view(np.uint8))
    # the same samples at a base address that is only texel-aligned take the scalar-load variant: same result
    flat = torch.zeros(ms.size + 4, dtype=src.dtype, device=cuda)
    off = flat[4:].view(ms.shape)  # +4 elements = one texel: 16-byte alignment is lost for 4- and 8-byte texels
    off.copy_(src)
    dst.zero_()
    ovr.resolve_msaa(off, dst, samples, fmt=fmt)
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy().view(np.uint8), want.view(np.uint8))
    # odd sample counts (generic path), 3 of the samples
    if samples == 4:
        ms3 = np.ascontiguousarray(ms.reshape(h, w, samples, 4)[:, :, :3].reshape(h, w * 3, 4))
        dst.zero_()
        ovr.resolve_msaa(torch.from_numpy(ms3).to(cuda), dst, 3, fmt=fmt)
        torch.cuda.synchronize()
        assert np.array_equal(dst.cpu().numpy().view(np.uint8), fe.resolve_msaa(ms3, 3, fmt).view(np.uint8))


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_apply_resolves_multisampled_sources_first(cuda, mode):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import frontend as fe, pyoracle as po
    iw, ih, samples = 233, 141, 4
    ms = _msaa(synth.natural_rgba8(iw, ih, 7), samples, 3, 0)
    resolved = fe.resolve_msaa(ms, samples, fe.FMT_RGBA8)
    ow, oh = po.output_size(iw, ih, 0.75)
    math = ovr.MATH_STRICT if mode == "strict" else ovr.MATH_FAST
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5, mathMode=math))
    out = pp.apply(0, torch.from_numpy(ms).to(cuda), samples=samples)
    torch.cuda.synchronize()
    mid = po.easu(resolved, ow, oh, po.upscale_constants(0, True, iw, ih, ow, oh, radius=0.5))
    want = po.rcas(mid, po.sharpen_constants(0, True, ow, oh, radius=0.5, sharpness=0.9))
    got = out.cpu().numpy()
    if mode == "strict":
        assert np.array_equal(got, want)
    else:  # the resolve is the same arithmetic in both modes; the passes are within 1 LSB each (amplified by RCAS)
        assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 4
    pp.close()
    # the stateless dispatches refuse multisampled images
    dst = torch.zeros((oh, ow, 4), dtype=torch.uint8, device=cuda)
    from openvr_fsr_b200 import _lib as L
    import ctypes as C
    s, d = ovr.image_of(torch.from_numpy(ms).to(cuda), None, samples), ovr.image_of(dst)
    uc = (C.c_uint32 * 24)(*po.upscale_constants(0, True, iw, ih, ow, oh).words())
    assert L.lib().ovrfsr_dispatch_fsr_easu(C.byref(s), C.byref(d), uc, L.MATH_STRICT, None) == L.ERR_UNSUPPORTED


def test_f7_capture_writes_the_left_eye_output(cuda, tmp_path):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import frontend as fe
    img = synth.natural_rgba8(200, 120, 4)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.6))
    pp.request_capture(str(tmp_path))
    right = pp.apply(1, ovr.to_image(img, cuda)).cpu().numpy()
    assert pp.last_capture_path() == "" and not list(tmp_path.iterdir())  # only Eye_Left triggers the capture
    left = pp.apply(0, ovr.to_image(img, cuda)).cpu().numpy()
    path = pp.last_capture_path()
    assert path.startswith(str(tmp_path)) and path.endswith("_fsr_s90_r60.dds")
    with open(path, "rb") as f:
        assert f.read() == fe.dds_bytes(left, fe.FMT_RGBA8)
    back, fmt = ovr.load_dds(path)
    assert fmt == ovr.FORMAT_RGBA8 and np.array_equal(back, left)
    # one shot: the next frame does not write again
    pp.apply(0, ovr.to_image(img, cuda))
    assert len(list(tmp_path.iterdir())) == 1
    pp.close()
    # a 10-bit chain captures a DX10-extension file
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75))
    pp.request_capture(str(tmp_path))
    out10 = pp.apply(0, ovr.to_image(synth.natural_rgb10a2(90, 60, 2), cuda), fmt=ovr.FORMAT_RGB10A2).cpu().numpy()
    with open(pp.last_capture_path(), "rb") as f:
        assert f.read() == fe.dds_bytes(out10, fe.FMT_RGB10A2)
    pp.close()


@pytest.mark.parametrize("samples", [2, 4, 8])
def test_resolve_is_the_exact_sample_mean_within_unorm_rounding(cuda, samples):
    """An independent statement of ResolveSubresource for UNORM targets (D3D11 functional spec: the resolved value is
    the average of the samples, written with the FLOAT -> UNORM conversion whose tolerance is 0.6 ULP): computed here in
    exact integer arithmetic, not through oracle/frontend.py.  |code - exact mean| <= 0.5 (+ float slack) everywhere,
    and equal to round-half-up of the exact mean wherever the mean is not within 1e-3 of a .5 tie."""
    import torch
    import openvr_fsr_b200 as ovr
    w, h = 67, 41
    rng = np.random.default_rng(samples)
    src = rng.integers(0, 256, (h, w * samples, 4), dtype=np.uint8)
    src[:, : 8 * samples] = np.repeat(rng.integers(0, 256, (h, 8, 4), dtype=np.uint8), samples, axis=1)  # some flat texels
    dst = torch.zeros((h, w, 4), dtype=torch.uint8, device=cuda)
    ovr.resolve_msaa(torch.from_numpy(src).to(cuda), dst, samples)
    torch.cuda.synchronize()
    got = dst.cpu().numpy().astype(np.float64)
    exact = src.reshape(h, w, samples, 4).astype(np.int64).sum(axis=2) / samples
    assert np.abs(got - exact).max() <= 0.5 + 1e-3
    frac = exact - np.floor(exact)
    clear = np.abs(frac - 0.5) > 1e-3
    assert np.array_equal(got[clear], np.floor(exact + 0.5)[clear])
    assert np.array_equal(got[:, :8], src[:, : 8 * samples : samples].astype(np.float64))  # identical samples resolve to themselves
