"""GPU parity, proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

strict math mode -> bit-identical to the oracle (which is itself bit-identical to the reference's own lines
compiled on the host, tests/test_oracle_vs_ref.py); fast mode -> <= 1 LSB per RGBA8 channel (north_star).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [  # (inW, inH, renderScale)
    (17, 13, 0.75), (16, 16, 0.5), (33, 47, 0.75), (200, 150, 0.77), (129, 65, 0.59), (64, 64, 1.3), (301, 97, 0.67),
]


def _inputs(w, h):
    from openvr_fsr_b200 import synth
    yield "uniform", synth.uniform_rgba8(w, h, 0)
    yield "natural", synth.natural_rgba8(w, h, 1)


def _run_gpu_easu(cuda, src_np, ow, oh, consts, mode, out_dtype=np.uint8):
    import torch
    import openvr_fsr_b200 as ovr
    src = torch.from_numpy(src_np).to(cuda)
    dst = torch.zeros((oh, ow, 4), dtype=torch.uint8 if out_dtype == np.uint8 else torch.float16, device=cuda)
    ovr.fsr_easu(src, dst, consts, mode)
    torch.cuda.synchronize()
    return dst.cpu().numpy()


def _run_gpu_rcas(cuda, src_np, consts, mode, out_dtype=np.uint8):
    import torch
    import openvr_fsr_b200 as ovr
    src = torch.from_numpy(src_np).to(cuda)
    dst = torch.zeros(src_np.shape, dtype=torch.uint8 if out_dtype == np.uint8 else torch.float16, device=cuda)
    ovr.fsr_rcas(src, dst, consts, mode)
    torch.cuda.synchronize()
    return dst.cpu().numpy()


@pytest.mark.parametrize("iw,ih,scale", SIZES)
@pytest.mark.parametrize("radius", [2.0, 0.5])
def test_easu_rcas_vs_oracle(cuda, iw, ih, scale, radius):
    import openvr_fsr_b200 as ovr
    from oracle import pyoracle as po
    ow, oh = po.output_size(iw, ih, scale)
    assert (ow, oh) == ovr.output_size(iw, ih, scale)
    uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=radius)
    sc = po.sharpen_constants(0, True, ow, oh, radius=radius, sharpness=0.9)
    for name, src in _inputs(iw, ih):
        ref_e = po.easu(src, ow, oh, uc)
        ref_r = po.rcas(ref_e, sc)
        # strict: bit-identical
        got_e = _run_gpu_easu(cuda, src, ow, oh, uc.words(), ovr.MATH_STRICT)
        assert np.array_equal(got_e, ref_e), f"EASU strict mismatch on {name}: {(got_e != ref_e).sum()} bytes"
        got_r = _run_gpu_rcas(cuda, ref_e, sc.words(), ovr.MATH_STRICT)
        assert np.array_equal(got_r, ref_r), f"RCAS strict mismatch on {name}: {(got_r != ref_r).sum()} bytes"
        # fast: <= 1 LSB per channel
        fast_e = _run_gpu_easu(cuda, src, ow, oh, uc.words(), ovr.MATH_FAST)
        assert np.abs(fast_e.astype(np.int16) - ref_e.astype(np.int16)).max() <= 1
        fast_r = _run_gpu_rcas(cuda, ref_e, sc.words(), ovr.MATH_FAST)
        assert np.abs(fast_r.astype(np.int16) - ref_r.astype(np.int16)).max() <= 1


def test_smoke_entry(cuda):
    import __graft_entry__ as g
    g.smoke()
