"""Dev tool: float-level comparison of the strict CUDA path vs the oracle (RGBA32F output)."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
from oracle import pyoracle as po

print("cpu denormal check:", np.float32(1e-39) * np.float32(1.0), np.float32(1e-20) * np.float32(1e-20))
dev = torch.device("cuda:0")
for (iw, ih, scale) in [(129, 65, 0.59), (200, 150, 0.77)]:
    ow, oh = po.output_size(iw, ih, scale)
    for radius in (2.0, 0.0):
        uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=radius)
        sc = po.sharpen_constants(0, True, ow, oh, radius=radius)
        for name, src in (("uniform", synth.uniform_rgba8(iw, ih, 0)), ("natural", synth.natural_rgba8(iw, ih, 1))):
            ref = po.easu(src, ow, oh, uc, out_dtype=np.float32)
            dst = torch.zeros((oh, ow, 4), dtype=torch.float32, device=dev)
            ovr.fsr_easu(torch.from_numpy(src).to(dev), dst, uc.words(), ovr.MATH_STRICT)
            torch.cuda.synchronize()
            got = dst.cpu().numpy()
            neq = got.view(np.uint32) != ref.view(np.uint32)
            px = neq.any(axis=2)
            d = np.abs(got - ref)
            print(f"EASU {iw}x{ih} r={radius} {name}: float-differing values {neq.sum()} in {px.sum()} px of {ow*oh}; "
                  f"max abs diff {d.max():.3e}; px with all 3 ch differing {int((neq[..., :3].sum(2) == 3).sum())}")
            ys, xs = np.nonzero(px)
            for y, x in list(zip(ys, xs))[:4]:
                print("    ", x, y, got[y, x, :3], ref[y, x, :3], "ulps", (got[y, x, :3].view(np.int32) - ref[y, x, :3].view(np.int32)))
            # RCAS on the oracle's 8-bit EASU output
            e8 = po.easu(src, ow, oh, uc)
            rref = po.rcas(e8, sc, out_dtype=np.float32)
            rdst = torch.zeros((oh, ow, 4), dtype=torch.float32, device=dev)
            ovr.fsr_rcas(torch.from_numpy(e8).to(dev), rdst, sc.words(), ovr.MATH_STRICT)
            torch.cuda.synchronize()
            rg = rdst.cpu().numpy()
            rneq = rg.view(np.uint32) != rref.view(np.uint32)
            print(f"RCAS: float-differing values {rneq.sum()}, max abs {np.abs(rg - rref).max():.3e}")
