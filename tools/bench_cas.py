"""Dev bench: the legacy CAS kernels at C2 (upscale 1683x1869 -> 2244x2492, sharpen at 2244x2492), per launch."""
import json, os, statistics, sys
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
ku, ks = ovr.cas_setup(0.9, 1.0, IW, IH, OW, OH), ovr.cas_setup(0.9, 1.0, OW, OH, OW, OH)
UP_B, SH_B = IW * IH * 4 + OW * OH * 4, 2 * OW * OH * 4
out = {}
for math, name in ((ovr.MATH_STRICT, "strict"), (ovr.MATH_FAST, "fast")):
    marks = []
    for rep in range(6):
        for i in range(8):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(); ovr.cas(pool[i], mid[i], ku, False, math); e[1].record()
            ovr.cas(mid[i], dst, ks, True, math); e[2].record()
            marks.append(e)
    torch.cuda.synchronize()
    tu = statistics.mean(m[0].elapsed_time(m[1]) for m in marks[8:]) * 1e3
    ts = statistics.mean(m[1].elapsed_time(m[2]) for m in marks[8:]) * 1e3
    out[name] = {"upscale_us": round(tu, 1), "upscale_GBps": round(UP_B / tu / 1e3), "sharpen_us": round(ts, 1),
                 "sharpen_GBps": round(SH_B / ts / 1e3)}
    print(name, out[name])
json.dump({"workload": "C2 eye, RGBA8", "algorithmic_bytes": {"upscale": UP_B, "sharpen": SH_B}, "results": out},
          open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "cas_bench.json"), "w"), indent=1)
