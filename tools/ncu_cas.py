"""Dev: a few CAS upscale / sharpen launches at the C2 eye size for ncu (strict math)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
dev = torch.device("cuda:0")
IW, IH, OW, OH = 1683, 1869, 2244, 2492
base = synth.natural_rgba8(IW, IH, 1)
pool = [ovr.to_image(np.roll(base, 37 * i, axis=0), dev) for i in range(6)]
mid = ovr.alloc_image(OW, OH, torch.uint8, dev); dst = ovr.alloc_image(OW, OH, torch.uint8, dev)
ku, ks = ovr.cas_setup(0.9, 1.0, IW, IH, OW, OH), ovr.cas_setup(0.9, 1.0, OW, OH, OW, OH)
for p in pool:
    ovr.cas(p, mid, ku, False, ovr.MATH_STRICT)
    ovr.cas(mid, dst, ks, True, ovr.MATH_STRICT)
torch.cuda.synchronize()
