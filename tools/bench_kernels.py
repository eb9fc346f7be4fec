"""Dev bench: per-launch time of the stateless EASU / RCAS / NVScaler dispatches at C2 / C4 (CUDA events, back to back on
one stream).  OVRFSR_LIB=<path> measures another build of the library on the same box."""
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
def timeit(fn, n=6):
    marks = []
    for rep in range(n):
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(i); e1.record(); marks.append((e0, e1))
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in marks[8:]) * 1e3
out = []
for math, mname in ((ovr.MATH_STRICT, "strict"), (ovr.MATH_FAST, "fast")):
    for radius in (2.0, 0.5, 0.0):
        cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=radius, mathMode=math,
                         projCentre=(0.5, 0.5, 0.5, 0.5) if radius > 0 else (9.0, 9.0, 9.0, 9.0))
        uc = ovr.make_upscale_constants(cfg, 0, True, IW, IH, OW, OH)
        sc = ovr.make_sharpen_constants(cfg, 0, True, OW, OH)
        te = timeit(lambda i: ovr.fsr_easu(pool[i], mid[i], uc, math))
        tr = timeit(lambda i: ovr.fsr_rcas(mid[i], dst, sc, math))
        out.append(f"{mname} r{radius}: EASU {te:6.1f} RCAS {tr:6.1f}")
iw, ih = 1512, 1680
ow, oh = ovr.output_size(iw, ih, 0.75)
npool = [ovr.to_image(np.roll(synth.natural_rgba8(iw, ih, 1), 31 * i, axis=0), dev) for i in range(8)]
ndst = ovr.alloc_image(ow, oh, torch.uint8, dev)
for math, mname in ((ovr.MATH_STRICT, "strict"), (ovr.MATH_FAST, "fast")):
    for radius in (2.0, 0.5):
        cfg = ovr.Config(fsrEnabled=True, useNis=True, renderScale=0.75, sharpness=0.9, radius=radius)
        scfg, _ = ovr.make_nis_config(cfg, False, 0, True, iw, ih, ow, oh)
        tn = timeit(lambda i: ovr.nis_scaler(npool[i], ndst, scfg, math))
        out.append(f"{mname} r{radius}: NVScaler {tn:6.1f}")
print(os.environ.get("OVRFSR_LIB", "current"), " | ".join(out))
# C3b: RCAS only on a 2915x3240 RGBA16F frame -> RGBA8 (plain-load tile path), radius 0.5
cw, ch = 2915, 3240
f16 = [torch.from_numpy(np.roll(synth.natural_rgba16f(cw, ch, 3), 37 * i, axis=0)).to(dev) for i in range(4)]
cdst = ovr.alloc_image(cw, ch, torch.uint8, dev)
o2 = []
for math, mname in ((ovr.MATH_STRICT, "strict"), (ovr.MATH_FAST, "fast")):
    cfg = ovr.Config(fsrEnabled=True, renderScale=1.0, sharpness=0.9, radius=0.5, mathMode=math)
    sc = ovr.make_sharpen_constants(cfg, 0, True, cw, ch)
    marks = []
    for rep in range(8):
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ovr.fsr_rcas(f16[i], cdst, sc, math); e1.record(); marks.append((e0, e1))
    torch.cuda.synchronize()
    o2.append(f"{mname} C3b RCAS fp16 {statistics.mean(a.elapsed_time(b) for a, b in marks[4:]) * 1e3:6.1f}")
print(" | ".join(o2))
