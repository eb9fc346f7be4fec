import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
from oracle import pyoracle as po
dev = torch.device("cuda:0")
for (iw, ih, scale) in [(16, 16, 0.5), (64, 48, 0.75)]:
    ow, oh = po.output_size(iw, ih, scale)
    uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=2.0)
    sc = po.sharpen_constants(0, True, ow, oh, radius=2.0)
    src = synth.uniform_rgba8(iw, ih, 0)
    t = torch.from_numpy(src).to(dev)
    for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
        print("rcas", iw, ih, mode, flush=True)
        e = po.easu(src, ow, oh, uc)
        d = torch.zeros((oh, ow, 4), dtype=torch.uint8, device=dev)
        ovr.fsr_rcas(torch.from_numpy(e).to(dev), d, sc.words(), mode); torch.cuda.synchronize()
        print("  rcas maxdiff", np.abs(d.cpu().numpy().astype(int) - po.rcas(e, sc).astype(int)).max(), flush=True)
        print("easu", iw, ih, mode, flush=True)
        d = torch.zeros((oh, ow, 4), dtype=torch.uint8, device=dev)
        ovr.fsr_easu(t, d, uc.words(), mode); torch.cuda.synchronize()
        print("  easu maxdiff", np.abs(d.cpu().numpy().astype(int) - e.astype(int)).max(), flush=True)
