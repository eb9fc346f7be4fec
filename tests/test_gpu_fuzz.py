"""Seeded random parity sweep through the stateful path (ovrfsr_apply): random sizes, scales, radii, projection
centres, sharpness, formats, debug tint, FSR / NIS, one-eye / shared-texture submits -- strict math, bit-identical
to the oracle.  Also the extreme aspect ratios the index arithmetic has to survive."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_chain(po, img, fmt, cfg, eye, one_eye, ow, oh):
    ih, iw = img.shape[:2]
    kw = dict(proj=cfg["proj"], radius=cfg["radius"], debug=cfg["debug"])
    ten = fmt == po.FMT_RGB10A2
    dfmt = po.FMT_RGB10A2 if ten else None
    if cfg["nis"]:
        ncfg, _ = po.nis_config(cfg["scale"] == 1.0, eye, one_eye, iw, ih, ow, oh, sharpness=cfg["sharp"], **kw)
        fn = po.nis_sharpen if cfg["scale"] == 1.0 else po.nis_scaler
        return fn(img, ncfg, src_fmt=fmt, dst_fmt=dfmt) if cfg["scale"] == 1.0 else fn(img, ow, oh, ncfg, src_fmt=fmt, dst_fmt=dfmt)
    mid, mfmt = img, fmt
    if cfg["scale"] != 1.0:
        mid = po.easu(img, ow, oh, po.upscale_constants(eye, one_eye, iw, ih, ow, oh, **{k: kw[k] for k in ("proj", "radius")}),
                      src_fmt=fmt, dst_fmt=dfmt)
        mfmt = dfmt
    return po.rcas(mid, po.sharpen_constants(eye, one_eye, ow, oh, sharpness=cfg["sharp"], **kw), src_fmt=mfmt, dst_fmt=dfmt)


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration(cuda, seed):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    rng = np.random.default_rng(1000 + seed)
    nis = bool(rng.integers(0, 2)) and seed % 3 == 0
    scale = float(rng.choice([0.5, 0.59, 0.67, 0.75, 0.77, 0.83, 0.91, 1.0] + ([] if nis else [1.3, 1.5])))
    iw, ih = int(rng.integers(9, 260)), int(rng.integers(9, 200))
    cfg = dict(nis=nis, scale=scale, sharp=float(rng.uniform(0.0, 1.0)), radius=float(rng.choice([0.0, 0.2, 0.45, 0.7, 2.0])),
               debug=bool(rng.integers(0, 2)), proj=tuple(float(x) for x in rng.uniform(0.3, 0.7, 4)))
    fmt = int(rng.choice([po.FMT_RGBA8, po.FMT_BGRA8, po.FMT_RGBA16F, po.FMT_RGB10A2]))
    one_eye = bool(rng.integers(0, 4))  # mostly one eye per texture; sometimes both side by side
    img = {po.FMT_RGBA8: synth.natural_rgba8, po.FMT_BGRA8: synth.uniform_rgba8, po.FMT_RGBA16F: synth.natural_rgba16f,
           po.FMT_RGB10A2: synth.natural_rgb10a2}[fmt](iw, ih, seed)
    ow, oh = po.output_size(iw, ih, scale)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, useNis=nis, renderScale=scale, sharpness=cfg["sharp"], radius=cfg["radius"],
                                      debugMode=cfg["debug"], projCentre=cfg["proj"]))
    bounds = ovr.TextureBounds() if one_eye else ovr.TextureBounds(0.0, 0.0, 0.5, 1.0)
    tex = ovr.to_image(img, cuda)
    for eye in ((0, 1) if one_eye else (0,)):
        out = pp.apply(eye, tex, bounds, fmt=None if fmt == po.FMT_RGBA16F else fmt)
        torch.cuda.synchronize()
        want = _oracle_chain(po, img, fmt, cfg, eye, one_eye, ow, oh)
        got = out.cpu().numpy()
        assert got.shape == want.shape
        assert np.array_equal(got, want), (cfg, fmt, iw, ih, eye, int((got != want).sum()))
    pp.close()


@pytest.mark.parametrize("iw,ih", [(16384, 9), (9, 16384), (8193, 33)])
def test_extreme_aspect_ratios(cuda, iw, ih):
    """D3D11's maximum texture dimension is 16384: tile counts, TMA box origins and row offsets at the limit."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    img = synth.uniform_rgba8(iw, ih, 3)
    scale = 1.0 if max(iw, ih) == 16384 else 0.75  # 16384 is already the largest target: sharpen only
    ow, oh = po.output_size(iw, ih, scale)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=2.0))
    out = pp.apply(0, ovr.to_image(img, cuda))
    torch.cuda.synchronize()
    mid = img if scale == 1.0 else po.easu(img, ow, oh, po.upscale_constants(0, True, iw, ih, ow, oh, radius=2.0), nthreads=8)
    want = po.rcas(mid, po.sharpen_constants(0, True, ow, oh, radius=2.0, sharpness=0.9), nthreads=8)
    assert np.array_equal(out.cpu().numpy(), want)
    pp.close()
    if scale == 1.0:  # and upscaling INTO the limit
        sw, sh = (12288, 9) if iw > ih else (9, 12288)
        src = synth.uniform_rgba8(sw, sh, 4)
        ow, oh = po.output_size(sw, sh, 0.75)
        assert max(ow, oh) == 16384
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=2.0))
        out = pp.apply(0, ovr.to_image(src, cuda))
        torch.cuda.synchronize()
        mid = po.easu(src, ow, oh, po.upscale_constants(0, True, sw, sh, ow, oh, radius=2.0), nthreads=8)
        assert np.array_equal(out.cpu().numpy(), po.rcas(mid, po.sharpen_constants(0, True, ow, oh, radius=2.0, sharpness=0.9), nthreads=8))
        pp.close()
