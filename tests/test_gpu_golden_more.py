"""The CUDA path (strict math, through the C ABI) against the committed single-pass fixtures of the NIS, CAS and
R10G10B10A2 paths -- outputs of the reference's own lines (tests/golden/make_golden_more.py)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = sorted((Path(__file__).parent / "golden").glob("pass_*.npz"))


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_cuda_matches_pass_golden(cuda, path):
    import torch
    import openvr_fsr_b200 as ovr
    g = np.load(path)
    kind, src_np, want = str(g["kind"]), g["src"], g["out"]
    sf, df = int(g["src_fmt"]), int(g["dst_fmt"])
    src = ovr.to_image(src_np, cuda)
    dst = torch.zeros(want.shape, dtype={np.dtype(np.uint8): torch.uint8, np.dtype(np.float16): torch.float16,
                                          np.dtype(np.float32): torch.float32}[want.dtype], device=cuda)
    kw = dict(src_fmt=None if sf < 0 else sf, dst_fmt=None if df < 0 else df)
    words = g["consts"].astype(np.uint32)
    for mode, tol in ((ovr.MATH_STRICT, 0), (ovr.MATH_FAST, 1)):
        dst.zero_()
        if kind == "nis_scaler":
            ovr.nis_scaler(src, dst, words.tobytes(), mode, **kw)
        elif kind == "nis_sharpen":
            ovr.nis_sharpen(src, dst, words.tobytes(), mode, **kw)
        elif kind == "cas":
            ovr.cas(src, dst, words, bool(int(g["sharpen_only"])), mode, **kw)
        elif kind == "easu":
            ovr.fsr_easu(src, dst, words, mode, **kw)
        elif kind == "rcas":
            ovr.fsr_rcas(src, dst, words, mode, **kw)
        torch.cuda.synchronize()
        got = dst.cpu().numpy()
        if tol == 0:
            assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), f"{kind}: strict differs from the reference lines"
        elif want.dtype == np.uint8 and df != ovr.FORMAT_RGB10A2:
            assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1
        elif want.dtype == np.uint8:
            from openvr_fsr_b200 import synth
            assert np.abs(synth.unpack_rgb10a2(got) - synth.unpack_rgb10a2(want)).max() <= 1
        else:
            assert np.abs(got.astype(np.float32) - want.astype(np.float32)).max() <= 4e-3
