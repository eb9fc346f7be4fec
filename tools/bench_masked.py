"""Dev bench: per-kernel time of EASU and RCAS at C2 for several radii (2.0 = every group filtered, 0.5 = the
reference default, 0.0 = every group takes the bilinear / copy path), both math modes."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
dev = torch.device("cuda:0")
IW, IH, OW, OH = 1683, 1869, 2244, 2492
base = synth.natural_rgba8(IW, IH, 1)
pool = [ovr.to_image(np.roll(base, 37 * i, axis=0), dev) for i in range(8)]
mid = [ovr.alloc_image(OW, OH, torch.uint8, dev) for _ in range(8)]
dst = ovr.alloc_image(OW, OH, torch.uint8, dev)
EASU_B, RCAS_B = IW * IH * 4 + OW * OH * 4, 2 * OW * OH * 4
for math, mname in ((ovr.MATH_STRICT, "strict"), (ovr.MATH_FAST, "fast")):
    for radius in (2.0, 0.5, 0.0):
        cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=radius, mathMode=math,
                         projCentre=(0.5, 0.5, 0.5, 0.5) if radius > 0 else (9.0, 9.0, 9.0, 9.0))
        uc = ovr.make_upscale_constants(cfg, 0, True, IW, IH, OW, OH)
        sc = ovr.make_sharpen_constants(cfg, 0, True, OW, OH)
        te, tr = [], []
        marks = []
        for rep in range(6):
            for i in range(8):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record(); ovr.fsr_easu(pool[i], mid[i], uc, math); e[1].record()
                ovr.fsr_rcas(mid[i], dst, sc, math); e[2].record()
                marks.append(e)
        torch.cuda.synchronize()
        te = statistics.mean(m[0].elapsed_time(m[1]) for m in marks[8:]) * 1e3
        tr = statistics.mean(m[1].elapsed_time(m[2]) for m in marks[8:]) * 1e3
        fm = []
        for rep in range(6):
            for i in range(8):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                e[0].record(); ovr.fsr_fused(pool[i], dst, uc, sc, math); e[1].record()
                fm.append(e)
        torch.cuda.synchronize()
        tf = statistics.mean(m[0].elapsed_time(m[1]) for m in fm[8:]) * 1e3
        print(f"{mname} radius {radius}: FUSED {tf:6.1f} us ({EASU_B / tf / 1e3:6.0f} GB/s of its own {EASU_B} B)   two-pass sum {te + tr:6.1f} us")
        print(f"{mname} radius {radius}: EASU {te:6.1f} us ({EASU_B / te / 1e3:6.0f} GB/s)  RCAS {tr:6.1f} us ({RCAS_B / tr / 1e3:6.0f} GB/s)")
