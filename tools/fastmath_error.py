"""Dev/measurement tool: how far the FAST math mode's composed EASU -> RGBA8 -> RCAS result is from the strict one (which
is bit-identical to the reference lines), per image class, at the C2 size.  Prints the histogram of |difference| in LSB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
dev = torch.device("cuda:0")
IW, IH = 1683, 1869
for name, img in (("natural", synth.natural_rgba8(IW, IH, 1)), ("natural2", synth.natural_rgba8(IW, IH, 7)), ("uniform-noise", synth.uniform_rgba8(IW, IH, 0))):
    for radius in (2.0, 0.5):
        outs = {}
        for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
            pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=radius, mathMode=mode))
            outs[mode] = pp.apply(0, ovr.to_image(img, dev)).cpu().numpy().astype(np.int16)
            pp.close()
        d = np.abs(outs[ovr.MATH_FAST] - outs[ovr.MATH_STRICT])[..., :3]
        hist = np.bincount(d.ravel(), minlength=4)
        print(f"{name:14s} radius {radius}: max {d.max():3d} LSB; fraction >0: {(d > 0).mean():.2e}, >1: {(d > 1).mean():.2e}, >2: {(d > 2).mean():.2e}; "
              f"histogram {hist[:8].tolist()}")
