"""Dev bench: the MSAA resolve front-end (ovrfsr_resolve_msaa) at the C2 eye size, against the HBM roofline.
Algorithmic bytes per launch = (samples + 1) * bpp * width * height (every sample read once, every texel written once)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openvr_fsr_b200 as ovr

dev = torch.device("cuda:0")
W, H = 1683, 1869
peak = 6570.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = []
for name, dt, fmt, bpp in (("RGBA8", torch.uint8, None, 4), ("RGBA16F", torch.float16, None, 8)):
    for samples in (2, 4, 8):
        pool = [torch.randint(0, 200, (H, W * samples, 4), device=dev).to(dt) for _ in range(6)]  # > L2 in total
        dst = torch.zeros((H, W, 4), dtype=dt, device=dev)
        for p in pool:
            ovr.resolve_msaa(p, dst, samples)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            for p in pool:
                ovr.resolve_msaa(p, dst, samples)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * len(pool))
        nbytes = (samples + 1) * bpp * W * H
        out.append({"format": name, "samples": samples, "us": round(us, 2), "bytes": nbytes,
                    "GBps": round(nbytes / us / 1e3, 1), "frac_of_hbm_peak": round(nbytes / us / 1e3 / peak, 3)})
        print(out[-1])
json.dump({"kernel": "resolve_msaa_kernel", "size": [W, H], "hbm_peak_gbs": peak, "results": out},
          open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "frontend_bench.json"), "w"), indent=1)
