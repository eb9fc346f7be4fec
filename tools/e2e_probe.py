"""Dev probe: where does the end-to-end (ovrfsr_apply_host) time go?  Host enqueue cost per call, and throughput
against the number of contexts in flight, radius (kernel time) and math mode."""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
dev = torch.device("cuda:0")
IW, IH, OW, OH = 1683, 1869, 2244, 2492
n = 8
g = torch.Generator().manual_seed(1)
h_in = [[torch.randint(0, 256, (IH, IW, 4), dtype=torch.uint8, generator=g).pin_memory() for _ in range(2)] for _ in range(n)]
h_out = [[torch.empty((OH, OW, 4), dtype=torch.uint8).pin_memory() for _ in range(2)] for _ in range(n)]

def run(nctx, radius, math, reps=5):
    cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=radius, mathMode=math)
    pps = [ovr.PostProcessor(cfg) for _ in range(nctx)]
    ss = [[torch.cuda.Stream(), torch.cuda.Stream()] for _ in range(nctx)]
    def step():
        for i in range(n):
            for eye in (0, 1):
                pps[i % nctx].apply_host(eye, h_in[i][eye], h_out[i][eye], stream=ss[i % nctx][eye])
    step(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for p in pps: p.close()
    return n * reps / (t2 - t0), (t1 - t0) / (n * reps * 2) * 1e6

for nctx in (1, 2, 4):
    for radius in (2.0, 0.05):
        for math in (ovr.MATH_STRICT, ovr.MATH_FAST):
            v, us = run(nctx, radius, math)
            print(f"ctx={nctx} radius={radius} math={math}: {v:.0f} pairs/s, host enqueue {us:.0f} us/call")
