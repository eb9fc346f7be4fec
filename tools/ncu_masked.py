"""Dev: a few EASU / RCAS launches at C2 with every group OUTSIDE the radius (pure bilinear / copy path), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
dev = torch.device("cuda:0")
IW, IH, OW, OH = 1683, 1869, 2244, 2492
base = synth.natural_rgba8(IW, IH, 1)
pool = [ovr.to_image(np.roll(base, 37 * i, axis=0), dev) for i in range(8)]
mid = ovr.alloc_image(OW, OH, torch.uint8, dev); dst = ovr.alloc_image(OW, OH, torch.uint8, dev)
cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.0, projCentre=(9.0, 9.0, 9.0, 9.0))
uc = ovr.make_upscale_constants(cfg, 0, True, IW, IH, OW, OH); sc = ovr.make_sharpen_constants(cfg, 0, True, OW, OH)
for i in range(8):
    ovr.fsr_easu(pool[i], mid, uc, ovr.MATH_STRICT); ovr.fsr_rcas(mid, dst, sc, ovr.MATH_STRICT)
torch.cuda.synchronize()
