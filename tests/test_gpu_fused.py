"""The fused EASU -> RCAS kernel (ovrfsr_dispatch_fsr_fused; PostProcessor.apply uses it with Config.fusedFsr)
against the two dispatches it replaces (PostProcessor.cpp:586-594) and against the oracle, through the C ABI.
Bit-identical in BOTH math modes: the fused kernel runs the same device functions on the same operands."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F8, FB, F10 = 0, 1, 4


def _consts(ovr, iw, ih, ow, oh, radius, sharp, debug, proj, eye=0):
    cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=sharp, radius=radius, debugMode=debug, projCentre=proj)
    return ovr.make_upscale_constants(cfg, eye, True, iw, ih, ow, oh), ovr.make_sharpen_constants(cfg, eye, True, ow, oh)


def _run_pair(ovr, torch, src, ow, oh, uc, sc, mode, src_fmt=None, mid_fmt=None):
    mid = ovr.alloc_image(ow, oh, torch.uint8, src.device)
    two = ovr.alloc_image(ow, oh, torch.uint8, src.device)
    one = ovr.alloc_image(ow, oh, torch.uint8, src.device)
    one.fill_(0x5a)
    ovr.fsr_easu(src, mid, uc, mode, src_fmt=src_fmt, dst_fmt=mid_fmt)
    ovr.fsr_rcas(mid, two, sc, mode, src_fmt=mid_fmt, dst_fmt=mid_fmt)
    ovr.fsr_fused(src, one, uc, sc, mode, src_fmt=src_fmt, dst_fmt=mid_fmt)
    torch.cuda.synchronize()
    return one.cpu().numpy(), two.cpu().numpy()


SIZES = [(17, 13, 0.75), (33, 47, 0.75), (211, 157, 0.75), (150, 110, 0.5), (97, 140, 0.59), (300, 200, 0.77), (90, 70, 1.3)]


@pytest.mark.parametrize("iw,ih,scale", SIZES)
@pytest.mark.parametrize("radius,debug,proj", [(2.0, False, (.5, .5, .5, .5)), (0.5, False, (.5, .5, .5, .5)),
                                               (0.33, True, (.41, .56, .6, .5)), (0.0, False, (9., 9., 9., 9.))])
def test_fused_equals_two_dispatches_and_oracle(cuda, iw, ih, scale, radius, debug, proj):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    ow, oh = po.output_size(iw, ih, scale)
    uc, sc = _consts(ovr, iw, ih, ow, oh, radius, 0.9, debug, proj)
    for name, img in (("natural", synth.natural_rgba8(iw, ih, 4)), ("uniform", synth.uniform_rgba8(iw, ih, 5))):
        src = ovr.to_image(img, cuda)
        for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
            one, two = _run_pair(ovr, torch, src, ow, oh, uc, sc, mode)
            assert np.array_equal(one, two), (name, mode, int((one != two).sum()))
        # strict: also the oracle's EASU -> RGBA8 -> RCAS
        one, _ = _run_pair(ovr, torch, src, ow, oh, uc, sc, ovr.MATH_STRICT)
        ouc = po.upscale_constants(0, True, iw, ih, ow, oh, proj=proj, radius=radius)
        osc = po.sharpen_constants(0, True, ow, oh, proj=proj, radius=radius, sharpness=0.9, debug=debug)
        assert np.array_equal(ouc.words(), uc) and np.array_equal(osc.words(), sc)
        want = po.rcas(po.easu(img, ow, oh, ouc), osc)
        assert np.array_equal(one, want), name


def test_fused_corner_images(cuda):
    import torch
    import openvr_fsr_b200 as ovr
    from tests.cases import corner_images
    iw, ih, scale = 70, 45, 0.67
    ow, oh = ovr.output_size(iw, ih, scale)
    uc, sc = _consts(ovr, iw, ih, ow, oh, 0.6, 0.7, False, (.5, .5, .5, .5))
    for name, img in corner_images(iw, ih).items():
        src = ovr.to_image(img, cuda)
        for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
            one, two = _run_pair(ovr, torch, src, ow, oh, uc, sc, mode)
            assert np.array_equal(one, two), (name, mode)


@pytest.mark.parametrize("fmt_in,fmt_mid", [(FB, F8), (F10, F10)])
def test_fused_other_unorm_sources(cuda, fmt_in, fmt_mid):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    iw, ih, scale = 130, 90, 0.75
    ow, oh = ovr.output_size(iw, ih, scale)
    uc, sc = _consts(ovr, iw, ih, ow, oh, 0.45, 0.9, True, (.5, .5, .5, .5))
    img = synth.uniform_rgba8(iw, ih, 9)  # for RGB10A2 the same bytes are read as packed 10:10:10:2 texels
    src = ovr.to_image(img, cuda)
    for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
        one, two = _run_pair(ovr, torch, src, ow, oh, uc, sc, mode, src_fmt=fmt_in, mid_fmt=fmt_mid)
        assert np.array_equal(one, two), mode


def test_fused_float_and_unaligned_sources(cuda):
    """RGBA16F source (plain-load tile path) and an RGBA8 source whose pitch is not 16-byte aligned (no TMA box)."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    iw, ih, scale = 123, 77, 0.75
    ow, oh = ovr.output_size(iw, ih, scale)
    uc, sc = _consts(ovr, iw, ih, ow, oh, 0.5, 0.9, False, (.5, .5, .5, .5))
    img = synth.natural_rgba8(iw, ih, 11)
    tight = torch.from_numpy(img).to(cuda)  # 123 * 4 = 492-byte rows
    half = (torch.from_numpy(img).to(cuda).float() * (2.0 / 255.0)).half()
    for src in (tight, half):
        for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
            one, two = _run_pair(ovr, torch, src, ow, oh, uc, sc, mode)
            assert np.array_equal(one, two), (src.dtype, mode)


def test_fused_rejects_mismatched_constants(cuda):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    iw, ih = 64, 48
    ow, oh = ovr.output_size(iw, ih, 0.75)
    uc, _ = _consts(ovr, iw, ih, ow, oh, 0.5, 0.9, False, (.5, .5, .5, .5))
    _, sc = _consts(ovr, iw, ih, ow, oh, 0.4, 0.9, False, (.5, .5, .5, .5))
    src = ovr.to_image(synth.natural_rgba8(iw, ih, 1), cuda)
    dst = ovr.alloc_image(ow, oh, torch.uint8, cuda)
    with pytest.raises(ovr.OvrFsrError) as e:
        ovr.fsr_fused(src, dst, uc, sc)
    assert e.value.status == ovr.ERR_UNSUPPORTED


def test_postprocessor_fused_equals_two_pass_at_c2(cuda):
    """BASELINE.json configs[1] full size, reference default radius and mask off, both eyes."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    iw, ih = 1683, 1869
    left, right = synth.stereo_pair("natural", iw, ih, 1)
    for radius in (0.5, 2.0):
        for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
            outs = []
            for fused in (True, False):
                pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=radius, mathMode=mode,
                                                  fusedFsr=fused))
                n0 = ovr.kernel_launches()
                o = [pp.apply(eye, ovr.to_image(img, cuda)).cpu().numpy() for eye, img in ((0, left), (1, right))]
                assert ovr.kernel_launches() - n0 == (2 if fused else 4)
                outs.append(o)
                pp.close()
            assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), (radius, mode)
