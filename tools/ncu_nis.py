"""Dev: a few NVScaler / NVSharpen launches at C4 for ncu.  usage: python tools/ncu_nis.py [strict|fast] [natural|textured|uniform]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
mode = ovr.MATH_FAST if (len(sys.argv) > 1 and sys.argv[1] == "fast") else ovr.MATH_STRICT
gen = {"natural": synth.natural_rgba8, "textured": synth.textured_rgba8, "uniform": synth.uniform_rgba8}[sys.argv[2] if len(sys.argv) > 2 else "natural"]
dev = torch.device("cuda:0")
iw, ih = 1512, 1680
ow, oh = ovr.output_size(iw, ih, 0.75)
cfg = ovr.Config(fsrEnabled=True, useNis=True, renderScale=0.75, sharpness=0.9, radius=2.0)
scfg, _ = ovr.make_nis_config(cfg, False, 0, True, iw, ih, ow, oh)
shcfg, _ = ovr.make_nis_config(cfg, True, 0, True, ow, oh, ow, oh)
pool = [ovr.to_image(np.roll(gen(iw, ih, 1), 31 * i, axis=0), dev) for i in range(6)]
big = ovr.to_image(gen(ow, oh, 2), dev)
dst = ovr.alloc_image(ow, oh, torch.uint8, dev)
for p in pool:
    ovr.nis_scaler(p, dst, scfg, mode)
    ovr.nis_sharpen(big, dst, shcfg, mode)
torch.cuda.synchronize()
