"""The CUDA path (through the C ABI) against the committed golden fixtures generated from the reference's own
lines: strict mode bit-identical, fast mode <= 1 LSB (RGBA8) / <= 1e-3 (FP16 out)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = sorted((Path(__file__).parent / "golden").glob("fsr_*.npz"))


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_cuda_matches_golden(cuda, path):
    import torch
    import openvr_fsr_b200 as ovr
    g = np.load(path)
    fmt = int(g["src_fmt"])
    src = torch.from_numpy(g["src"]).to(cuda)
    oh, ow = g["easu"].shape[:2]
    f16 = g["easu"].dtype == np.float16
    tdt = torch.float16 if f16 else torch.uint8
    for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
        easu = torch.zeros((oh, ow, 4), dtype=tdt, device=cuda)
        ovr.fsr_easu(src, easu, g["upscale"], mode, src_fmt=None if fmt < 0 else fmt)
        rcas = torch.zeros_like(easu)
        ovr.fsr_rcas(torch.from_numpy(g["easu"]).to(cuda), rcas, g["sharpen"], mode)
        torch.cuda.synchronize()
        for got, want in ((easu.cpu().numpy(), g["easu"]), (rcas.cpu().numpy(), g["rcas"])):
            if mode == ovr.MATH_STRICT:
                assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
            elif f16:
                assert np.abs(got.astype(np.float32) - want.astype(np.float32)).max() <= 1e-3 * max(1.0, float(np.abs(want.astype(np.float32)).max()))
            else:
                assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1
