"""GPU parity for DXGI_FORMAT_R10G10B10A2_UNORM eye textures: the format DetermineOutputFormat keeps 10-bit
(PostProcessor.cpp:63-74).  Strict math is bit-identical to the oracle; fast math is within 1 code (of 1023) per pass."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F10 = 4  # OVRFSR_FORMAT_RGB10A2 == oracle FMT_RGB10A2


def _codes(img):
    from openvr_fsr_b200 import synth
    return synth.unpack_rgb10a2(img)


def _gpu(cuda, fn, src_np, shape, consts, mode, dst_fmt, src_fmt=F10):
    import torch
    dt = {F10: torch.uint8, 2: torch.float16, 3: torch.float32}[dst_fmt]
    src = torch.from_numpy(src_np).to(cuda)
    dst = torch.zeros(shape, dtype=dt, device=cuda)
    fn(src, dst, consts, mode, src_fmt=src_fmt, dst_fmt=dst_fmt if dst_fmt == F10 else None)
    torch.cuda.synchronize()
    return dst.cpu().numpy()


@pytest.mark.parametrize("iw,ih,scale", [(17, 13, 0.75), (200, 150, 0.77), (301, 97, 0.67), (129, 65, 0.59), (333, 211, 0.75)])
@pytest.mark.parametrize("radius", [2.0, 0.45])
def test_fsr_rgb10a2_vs_oracle(cuda, iw, ih, scale, radius):
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    assert ovr.FORMAT_RGB10A2 == po.FMT_RGB10A2 == F10
    ow, oh = po.output_size(iw, ih, scale)
    uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=radius)
    sc = po.sharpen_constants(0, True, ow, oh, radius=radius, sharpness=0.9)
    for src in (synth.natural_rgb10a2(iw, ih, 3), synth.uniform_rgb10a2(iw, ih, 4)):
        ref_e = po.easu(src, ow, oh, uc, src_fmt=F10, dst_fmt=F10)
        ref_r = po.rcas(ref_e, sc, src_fmt=F10, dst_fmt=F10)
        got_e = _gpu(cuda, ovr.fsr_easu, src, (oh, ow, 4), uc.words(), ovr.MATH_STRICT, F10)
        assert np.array_equal(got_e, ref_e), f"EASU strict: {(got_e != ref_e).sum()} bytes differ"
        got_r = _gpu(cuda, ovr.fsr_rcas, ref_e, (oh, ow, 4), sc.words(), ovr.MATH_STRICT, F10)
        assert np.array_equal(got_r, ref_r), f"RCAS strict: {(got_r != ref_r).sum()} bytes differ"
        # float outputs of the same passes (pre-quantisation values)
        for dfmt, odt in ((3, np.float32), (2, np.float16)):
            want = po.easu(src, ow, oh, uc, out_dtype=odt, src_fmt=F10)
            got = _gpu(cuda, ovr.fsr_easu, src, (oh, ow, 4), uc.words(), ovr.MATH_STRICT, dfmt)
            assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
            want = po.rcas(ref_e, sc, out_dtype=odt, src_fmt=F10)
            got = _gpu(cuda, ovr.fsr_rcas, ref_e, (oh, ow, 4), sc.words(), ovr.MATH_STRICT, dfmt)
            assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
        fast_e = _gpu(cuda, ovr.fsr_easu, src, (oh, ow, 4), uc.words(), ovr.MATH_FAST, F10)
        assert np.abs(_codes(fast_e) - _codes(ref_e)).max() <= 1
        fast_r = _gpu(cuda, ovr.fsr_rcas, ref_e, (oh, ow, 4), sc.words(), ovr.MATH_FAST, F10)
        assert np.abs(_codes(fast_r) - _codes(ref_r)).max() <= 1


def test_rcas_outside_radius_passes_10bit_texels_through(cuda):
    """alpha included (fsr_rcas.hlsl:45-53): with every group outside the radius the output equals the input."""
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    src = synth.uniform_rgb10a2(150, 70, 8)
    sc = po.sharpen_constants(0, True, 150, 70, radius=0.0, sharpness=0.9, proj=(5.0, 5.0, 5.0, 5.0))
    for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
        assert np.array_equal(_gpu(cuda, ovr.fsr_rcas, src, src.shape, sc.words(), mode, F10), src)


@pytest.mark.parametrize("scale", [0.75, 1.0])
def test_postprocessor_keeps_10bit_target(cuda, scale):
    """FORMAT_AUTO: an RGB10A2 source gets RGB10A2 intermediates and output; the whole chain is bit-identical."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih = 233, 141
    img = synth.natural_rgb10a2(iw, ih, 12)
    ow, oh = po.output_size(iw, ih, scale)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.8, radius=0.5))
    for eye in (0, 1):
        out = pp.apply(eye, ovr.to_image(img, cuda), fmt=F10)
        torch.cuda.synchronize()
        assert pp.last_output_format == F10
        sc = po.sharpen_constants(eye, True, ow, oh, radius=0.5, sharpness=0.8)
        mid = img if scale == 1.0 else po.easu(img, ow, oh, po.upscale_constants(eye, True, iw, ih, ow, oh, radius=0.5),
                                                src_fmt=F10, dst_fmt=F10)
        assert np.array_equal(out.cpu().numpy(), po.rcas(mid, sc, src_fmt=F10, dst_fmt=F10))
    pp.close()
    # an 8-bit UNORM target is refused for a 10-bit source, and vice versa (DetermineOutputFormat)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, outputFormat=ovr.FORMAT_RGBA8))
    with pytest.raises(ovr.OvrFsrError):
        pp.apply(0, ovr.to_image(img, cuda), fmt=F10)
    pp.close()
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, outputFormat=F10))
    with pytest.raises(ovr.OvrFsrError):
        pp.apply(0, ovr.to_image(synth.natural_rgba8(iw, ih, 1), cuda))
    pp.close()


def test_nis_rgb10a2_vs_oracle(cuda):
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih, scale = 161, 95, 0.75
    ow, oh = po.output_size(iw, ih, scale)
    src = synth.natural_rgb10a2(iw, ih, 5)
    cfg, ok = po.nis_config(False, 0, True, iw, ih, ow, oh, radius=0.5, sharpness=0.7)
    assert ok
    want = po.nis_scaler(src, ow, oh, cfg, src_fmt=F10, dst_fmt=F10)
    got = _gpu(cuda, ovr.nis_scaler, src, (oh, ow, 4), bytes(cfg), ovr.MATH_STRICT, F10)
    assert np.array_equal(got, want)
    fast = _gpu(cuda, ovr.nis_scaler, src, (oh, ow, 4), bytes(cfg), ovr.MATH_FAST, F10)
    assert np.abs(_codes(fast) - _codes(want)).max() <= 1
    cfg, ok = po.nis_config(True, 0, True, iw, ih, iw, ih, radius=0.5, sharpness=0.7)
    want = po.nis_sharpen(src, cfg, src_fmt=F10, dst_fmt=F10)
    got = _gpu(cuda, ovr.nis_sharpen, src, (ih, iw, 4), bytes(cfg), ovr.MATH_STRICT, F10)
    assert np.array_equal(got, want)
    fast = _gpu(cuda, ovr.nis_sharpen, src, (ih, iw, 4), bytes(cfg), ovr.MATH_FAST, F10)
    assert np.abs(_codes(fast) - _codes(want)).max() <= 1
