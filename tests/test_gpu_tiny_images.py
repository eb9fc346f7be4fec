"""Degenerate image sizes (1x1, single rows / columns, a few texels): every tap is a border case -- clamp-to-edge,
zero Load, the TMA box lying mostly outside the image.  Strict math, bit-identical to the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [(1, 1), (2, 3), (5, 4), (3, 1), (1, 7), (8, 8)]


@pytest.mark.parametrize("iw,ih", SIZES)
@pytest.mark.parametrize("scale", [0.5, 0.75, 1.0])
def test_tiny_fsr_nis_cas(cuda, iw, ih, scale):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    img = synth.uniform_rgba8(iw, ih, 3 + iw * 7 + ih)
    ow, oh = po.output_size(iw, ih, scale)
    for aligned in (True, False):  # TMA box loads / plain loads
        tex = ovr.to_image(img, cuda) if aligned else torch.from_numpy(img).to(cuda)
        # FSR through the context
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=2.0))
        got = pp.apply(0, tex).cpu().numpy()
        pp.close()
        mid = img if scale == 1.0 else po.easu(img, ow, oh, po.upscale_constants(0, True, iw, ih, ow, oh, radius=2.0))
        assert np.array_equal(got, po.rcas(mid, po.sharpen_constants(0, True, ow, oh, radius=2.0, sharpness=0.9))), ("fsr", aligned)
        out = torch.zeros((oh, ow, 4), dtype=torch.uint8, device=cuda)
        # NIS
        ncfg, _ = po.nis_config(scale == 1.0, 0, True, iw, ih, ow, oh, radius=2.0, sharpness=0.8)
        if scale == 1.0:
            ovr.nis_sharpen(tex, out, bytes(ncfg), ovr.MATH_STRICT)
            want = po.nis_sharpen(img, ncfg)
        else:
            ovr.nis_scaler(tex, out, bytes(ncfg), ovr.MATH_STRICT)
            want = po.nis_scaler(img, ow, oh, ncfg)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want), ("nis", aligned)
        # CAS: the upscale shader at any size, the sharpen shader when sizes match
        kc = po.cas_setup(0.8, 1.0, iw, ih, ow, oh)
        ovr.cas(tex, out, kc.words(), False, ovr.MATH_STRICT)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), po.cas(img, ow, oh, kc, False)), ("cas upscale", aligned)
        if scale == 1.0:
            ovr.cas(tex, out, kc.words(), True, ovr.MATH_STRICT)
            torch.cuda.synchronize()
            assert np.array_equal(out.cpu().numpy(), po.cas(img, ow, oh, kc, True)), ("cas sharpen", aligned)
