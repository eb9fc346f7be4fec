"""GPU parity for the legacy CAS kernels (src/cas): strict math bit-identical to the oracle (itself
bit-identical to the reference's CasFilter lines, tests/test_cas.py), fast math within 1 LSB."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu(cuda, src_np, ow, oh, consts, sharpen_only, mode, out_dtype=np.uint8, src_fmt=None, aligned=True):
    import torch
    import openvr_fsr_b200 as ovr
    src = ovr.to_image(src_np, cuda) if aligned else torch.from_numpy(src_np).to(cuda)
    dt = {np.uint8: torch.uint8, np.float16: torch.float16, np.float32: torch.float32}[out_dtype]
    dst = torch.zeros((oh, ow, 4), dtype=dt, device=cuda)
    ovr.cas(src, dst, consts, sharpen_only, mode, src_fmt=src_fmt)
    torch.cuda.synchronize()
    return dst.cpu().numpy()


@pytest.mark.parametrize("w,h", [(17, 13), (64, 32), (200, 150), (301, 97), (1030, 70)])
@pytest.mark.parametrize("sharp,mcd", [(0.0, 1.0), (0.8, 1.0), (1.0, 0.08)])
def test_cas_sharpen_vs_oracle(cuda, w, h, sharp, mcd):
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    k = po.cas_setup(sharp, mcd, w, h, w, h)
    for src in (synth.natural_rgba8(w, h, 3), synth.uniform_rgba8(w, h, 4)):
        want = po.cas(src, w, h, k, True)
        for aligned in (True, False):  # TMA box loads / plain loads
            got = _gpu(cuda, src, w, h, k.words(), True, ovr.MATH_STRICT, aligned=aligned)
            assert np.array_equal(got, want), f"{(got != want).sum()} bytes differ (aligned={aligned})"
        fast = _gpu(cuda, src, w, h, k.words(), True, ovr.MATH_FAST)
        assert np.abs(fast.astype(np.int16) - want.astype(np.int16)).max() <= 1
        wantf = po.cas(src, w, h, k, True, out_dtype=np.float32)
        gotf = _gpu(cuda, src, w, h, k.words(), True, ovr.MATH_STRICT, out_dtype=np.float32)
        assert np.array_equal(gotf.view(np.uint8), wantf.view(np.uint8))
    # other source formats
    f16 = synth.natural_rgba16f(w, h, 5)
    assert np.array_equal(_gpu(cuda, f16, w, h, k.words(), True, ovr.MATH_STRICT, out_dtype=np.float16).view(np.uint8),
                          po.cas(f16, w, h, k, True, out_dtype=np.float16).view(np.uint8))
    bgra = synth.uniform_rgba8(w, h, 6)
    assert np.array_equal(_gpu(cuda, bgra, w, h, k.words(), True, ovr.MATH_STRICT, src_fmt=ovr.FORMAT_BGRA8),
                          po.cas(bgra, w, h, k, True, src_fmt=po.FMT_BGRA8))


@pytest.mark.parametrize("iw,ih,scale", [(17, 13, 0.75), (48, 40, 0.5), (200, 150, 0.77), (129, 65, 0.59), (64, 64, 1.0),
                                         (301, 97, 0.67), (240, 135, 0.75)])
@pytest.mark.parametrize("sharp", [0.0, 0.9])
def test_cas_upscale_vs_oracle(cuda, iw, ih, scale, sharp):
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    ow, oh = po.output_size(iw, ih, scale)
    k = po.cas_setup(sharp, 1.0, iw, ih, ow, oh)
    assert np.array_equal(ovr.cas_setup(sharp, 1.0, iw, ih, ow, oh), k.words())
    for src in (synth.natural_rgba8(iw, ih, 7), synth.uniform_rgba8(iw, ih, 8)):
        want = po.cas(src, ow, oh, k, False)
        got = _gpu(cuda, src, ow, oh, k.words(), False, ovr.MATH_STRICT)
        assert np.array_equal(got, want), f"{(got != want).sum()} bytes differ"
        fast = _gpu(cuda, src, ow, oh, k.words(), False, ovr.MATH_FAST)
        assert np.abs(fast.astype(np.int16) - want.astype(np.int16)).max() <= 1
        wantf = po.cas(src, ow, oh, k, False, out_dtype=np.float32)
        gotf = _gpu(cuda, src, ow, oh, k.words(), False, ovr.MATH_STRICT, out_dtype=np.float32)
        assert np.array_equal(gotf.view(np.uint8), wantf.view(np.uint8))
    f16 = synth.natural_rgba16f(iw, ih, 9)
    assert np.array_equal(_gpu(cuda, f16, ow, oh, k.words(), False, ovr.MATH_STRICT, out_dtype=np.float16).view(np.uint8),
                          po.cas(f16, ow, oh, k, False, out_dtype=np.float16).view(np.uint8))


def test_cas_c2_full_frame_and_shape_rules(cuda):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih = 1683, 1869
    ow, oh = po.output_size(iw, ih, 0.75)
    src = synth.natural_rgba8(iw, ih, 1)
    k = po.cas_setup(0.9, 1.0, iw, ih, ow, oh)
    want = po.cas(src, ow, oh, k, False, nthreads=8)
    assert np.array_equal(_gpu(cuda, src, ow, oh, k.words(), False, ovr.MATH_STRICT), want)
    ks = po.cas_setup(0.9, 1.0, ow, oh, ow, oh)
    assert np.array_equal(_gpu(cuda, want, ow, oh, ks.words(), True, ovr.MATH_STRICT), po.cas(want, ow, oh, ks, True, nthreads=8))
    # sharpen needs equal sizes; the upscale shader never shrinks
    t = torch.zeros((10, 10, 4), dtype=torch.uint8, device=cuda)
    with pytest.raises(ovr.OvrFsrError):
        ovr.cas(t, torch.zeros((12, 12, 4), dtype=torch.uint8, device=cuda), ks.words(), True)
    with pytest.raises(ovr.OvrFsrError):
        ovr.cas(t, torch.zeros((8, 8, 4), dtype=torch.uint8, device=cuda), ks.words(), False)
