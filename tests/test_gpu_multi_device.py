"""Two GPUs, one eye each: identical bits to the single-device result.  Skipped on a one-GPU box; the
world_size-2 host logic is covered on CPU by tests/test_sharding_gloo.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_one_eye_per_device_matches_oracle(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih, scale = 301, 211, 0.75
    left, right = synth.stereo_pair("natural", iw, ih, 5)
    ow, oh = po.output_size(iw, ih, scale)
    outs = {}
    for eye, img in ((0, left), (1, right)):
        with torch.cuda.device(eye):  # eye -> GPU, SURVEY 8e
            pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=0.5, device=eye))
            out = pp.apply(eye, ovr.to_image(img, torch.device("cuda", eye)))
            torch.cuda.synchronize(eye)
            outs[eye] = out.cpu().numpy()
            pp.close()
        mid = po.easu(img, ow, oh, po.upscale_constants(eye, True, iw, ih, ow, oh, radius=0.5))
        want = po.rcas(mid, po.sharpen_constants(eye, True, ow, oh, radius=0.5, sharpness=0.9))
        assert np.array_equal(outs[eye], want)
    # and the same eye processed on the other device gives the same bits (kernels are deterministic)
    with torch.cuda.device(1):
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=0.5, device=1))
        again = pp.apply(0, ovr.to_image(left, torch.device("cuda", 1))).cpu().numpy()
        pp.close()
    assert np.array_equal(again, outs[0])
    # every kernel that opts in to > 48 KB of dynamic shared memory must do so on EACH device it runs on
    # (regression: the opt-in used to be once per process): NVScaler and the CAS upscale on device 1, after device 0
    for d in (0, 1):
        with torch.cuda.device(d):
            dv = torch.device("cuda", d)
            src, out = ovr.to_image(left, dv), torch.zeros((oh, ow, 4), dtype=torch.uint8, device=dv)
            ncfg, _ = po.nis_config(False, 0, True, iw, ih, ow, oh, radius=0.5, sharpness=0.9)
            ovr.nis_scaler(src, out, bytes(ncfg), ovr.MATH_STRICT)
            torch.cuda.synchronize(d)
            assert np.array_equal(out.cpu().numpy(), po.nis_scaler(left, ow, oh, ncfg))
            kc = po.cas_setup(0.9, 1.0, iw, ih, ow, oh)
            ovr.cas(src, out, kc.words(), False, ovr.MATH_FAST)
            ovr.cas(src, out, kc.words(), False, ovr.MATH_STRICT)
            torch.cuda.synchronize(d)
            assert np.array_equal(out.cpu().numpy(), po.cas(left, ow, oh, kc, False))
