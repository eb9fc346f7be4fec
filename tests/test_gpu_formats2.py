"""B8G8R8X8 and R32G32B32_FLOAT sources and the _SRGB / _TYPELESS tags (PostProcessor.cpp:30-102) on the GPU path,
through the C ABI, against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_bgrx8_source_all_paths(cuda):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih, scale = 150, 97, 0.75
    ow, oh = po.output_size(iw, ih, scale)
    img = synth.natural_rgba8(iw, ih, 3)
    img[..., 3] = np.random.default_rng(1).integers(0, 256, (ih, iw), dtype=np.uint8)  # junk in the X byte
    for src in (ovr.to_image(img, cuda), torch.from_numpy(img).to(cuda)):  # TMA tile loads / plain loads (600-byte rows)
        # RCAS alone, small radius: the outside-radius copy passes source alpha through -> must read 1
        sc = po.sharpen_constants(0, True, iw, ih, radius=0.3, sharpness=0.8)
        out = ovr.alloc_image(iw, ih, torch.uint8, cuda)
        for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
            ovr.fsr_rcas(src, out, sc.words(), mode, src_fmt=ovr.FORMAT_BGRX8)
            torch.cuda.synchronize()
            want = po.rcas(img, sc, src_fmt=po.FMT_BGRX8)
            got = out.cpu().numpy()
            assert np.array_equal(got[..., 3], want[..., 3])
            if mode == ovr.MATH_STRICT:
                assert np.array_equal(got, want)
        # NVScaler / NVSharpen: alpha = bilinear of source alpha / source alpha
        cfg, _ = po.nis_config(False, 0, True, iw, ih, ow, oh, radius=2.0, sharpness=0.9)
        o2 = ovr.alloc_image(ow, oh, torch.uint8, cuda)
        ovr.nis_scaler(src, o2, bytes(cfg), ovr.MATH_STRICT, src_fmt=ovr.FORMAT_BGRX8)
        torch.cuda.synchronize()
        assert np.array_equal(o2.cpu().numpy(), po.nis_scaler(img, ow, oh, cfg, src_fmt=po.FMT_BGRX8))
        scfg, _ = po.nis_config(True, 0, True, iw, ih, iw, ih, radius=2.0, sharpness=0.9)
        ovr.nis_sharpen(src, out, bytes(scfg), ovr.MATH_STRICT, src_fmt=ovr.FORMAT_BGRX8)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), po.nis_sharpen(img, scfg, src_fmt=po.FMT_BGRX8))
    # the whole pass through the context, tagged as the sRGB variant (viewed as UNORM: same pixels)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=0.4))
    got = pp.apply(0, ovr.to_image(img, cuda), fmt=ovr.FORMAT_BGRX8 | ovr.FORMAT_SRGB_BIT).cpu().numpy()
    uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=0.4)
    sc = po.sharpen_constants(0, True, ow, oh, radius=0.4, sharpness=0.9)
    assert np.array_equal(got, po.rcas(po.easu(img, ow, oh, uc, src_fmt=po.FMT_BGRX8), sc))
    pp.close()


def test_rgb32f_source_through_the_context(cuda):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih, scale = 121, 90, 0.75
    ow, oh = po.output_size(iw, ih, scale)
    rgb = np.ascontiguousarray(synth.natural_rgba16f(iw, ih, 2).astype(np.float32)[..., :3])
    t = torch.from_numpy(rgb).to(cuda)
    uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=0.5)
    sc = po.sharpen_constants(0, True, ow, oh, radius=0.5, sharpness=0.9)
    want = po.rcas(po.easu(rgb, ow, oh, uc, src_fmt=po.FMT_RGB32F), sc)
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=0.5))
    got = pp.apply(0, t, fmt=ovr.FORMAT_RGB32F).cpu().numpy()
    assert np.array_equal(got, want)
    pp.close()
    # NIS path too
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, useNis=True, renderScale=scale, sharpness=0.9, radius=2.0))
    cfg, _ = po.nis_config(False, 0, True, iw, ih, ow, oh, radius=2.0, sharpness=0.9)
    assert np.array_equal(pp.apply(0, t, fmt=ovr.FORMAT_RGB32F).cpu().numpy(), po.nis_scaler(rgb, ow, oh, cfg, src_fmt=po.FMT_RGB32F))
    pp.close()
    # the expansion itself, and the stateless dispatches' refusal
    rgba = torch.zeros((ih, iw, 4), dtype=torch.float32, device=cuda)
    ovr.expand_rgb32f(t, rgba)
    torch.cuda.synchronize()
    r = rgba.cpu().numpy()
    assert np.array_equal(r[..., :3], rgb) and bool((r[..., 3] == 1.0).all())
    out = ovr.alloc_image(ow, oh, torch.uint8, cuda)
    with pytest.raises(ovr.OvrFsrError) as e:
        ovr.fsr_easu(t, out, uc.words(), ovr.MATH_STRICT, src_fmt=ovr.FORMAT_RGB32F)
    assert e.value.status == ovr.ERR_UNSUPPORTED


def test_srgb_and_typeless_tags_change_no_pixel(cuda):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    img = synth.natural_rgba8(96, 64, 5)
    outs = []
    for tag in (0, ovr.FORMAT_SRGB_BIT, ovr.FORMAT_TYPELESS_BIT):
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5))
        outs.append(pp.apply(0, ovr.to_image(img, cuda), fmt=ovr.FORMAT_RGBA8 | tag).cpu().numpy())
        pp.close()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
