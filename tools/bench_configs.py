"""Measurement tool: device-resident throughput of every BASELINE.json / SURVEY 8d configuration on ONE GPU, through the
stateful path (ovrfsr_apply), strict and fast math, two frames in flight with one stream per eye (as bench.py).
  C1  single eye 960x1080 -> 1920x2160 RGBA8, FSR, radius 2.0
  C2  stereo 1683x1869 -> 2244x2492 RGBA8, FSR, radius 2.0 and 0.5            (bench.py's workload)
  C3a stereo 2244x2492 RGBA16F -> 2917x3239 RGBA8, renderScale 1.3, EASU+RCAS, radius 0.5   (what the code does)
  C3b stereo 2915x3240 RGBA16F, RCAS only, radius 0.5, RGBA8 out                            (BASELINE.json's literal reading)
  C4  stereo 1512x1680 -> 2016x2240 RGBA8, NIS (NVScaler only), radius 2.0 and 0.5
Algorithmic bytes per eye as SURVEY 8d counts them."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth

dev = torch.device("cuda:0")


def run(cfg, images, reps=6):
    pps = [ovr.PostProcessor(cfg), ovr.PostProcessor(cfg)]
    streams = [[torch.cuda.Stream(), torch.cuda.Stream()] for _ in pps]
    main = torch.cuda.current_stream()

    def step():
        for i, (l, r) in enumerate(images):
            k = i & 1
            pps[k].apply(0, l, stream=streams[k][0])
            if r is not None:
                pps[k].apply(1, r, stream=streams[k][1])
    step(); step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for pair in streams:
        for s in pair: s.wait_stream(main)
    for _ in range(reps): step()
    for pair in streams:
        for s in pair: main.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    for p in pps: p.close()
    return len(images) * reps / (e0.elapsed_time(e1) * 1e-3)


def pool(make, w, h, n, stereo=True):
    base = make(w, h, 1)
    out = []
    for i in range(n):
        l = ovr.to_image(np.roll(base, 37 * i, axis=0), dev)
        out.append((l, ovr.to_image(np.roll(base, 37 * i + 16, axis=1), dev) if stereo else None))
    return out


results = []
def report(name, cfg_kw, images, bytes_per_unit, unit):
    for math, mname in ((ovr.MATH_STRICT, "strict"), (ovr.MATH_FAST, "fast")):
        v = run(ovr.Config(fsrEnabled=True, mathMode=math, **cfg_kw), images)
        results.append({"config": name, "math": mname, "value": round(v, 1), "unit": unit, "GBps_algorithmic": round(v * bytes_per_unit / 1e9, 1)})
        print(results[-1])


p = pool(synth.natural_rgba8, 960, 1080, 16, stereo=False)
report("C1 single eye 960x1080->1920x2160 FSR r2.0", dict(renderScale=0.5, sharpness=0.9, radius=2.0), p, 53913600, "eyes/s")
del p
p = pool(synth.natural_rgba8, 1683, 1869, 8)
report("C2 stereo 1683x1869->2244x2492 FSR r2.0", dict(renderScale=0.75, sharpness=0.9, radius=2.0), p, 159373368, "pairs/s")
report("C2 stereo 1683x1869->2244x2492 FSR r0.5", dict(renderScale=0.75, sharpness=0.9, radius=0.5), p, 159373368, "pairs/s")
del p
p = pool(synth.natural_rgba16f, 2244, 2492, 4)
report("C3a stereo 2244x2492 FP16 ->2917x3239 EASU+RCAS r0.5", dict(renderScale=1.3, sharpness=0.9, radius=0.5), p, 2 * 158114340, "pairs/s")
del p
p = pool(synth.natural_rgba16f, 2915, 3240, 4)
report("C3b stereo 2915x3240 FP16 RCAS only r0.5", dict(renderScale=1.0, sharpness=0.9, radius=0.5), p, 2 * 113335200, "pairs/s")
del p
p = pool(synth.natural_rgba8, 1512, 1680, 8)
report("C4 stereo 1512x1680->2016x2240 NIS r2.0", dict(useNis=True, renderScale=0.75, sharpness=0.9, radius=2.0), p, 2 * 28224000, "pairs/s")
report("C4 stereo 1512x1680->2016x2240 NIS r0.5", dict(useNis=True, renderScale=0.75, sharpness=0.9, radius=0.5), p, 2 * 28224000, "pairs/s")
# NVScaler skips the directional filters of edge-free pixels: the same configuration on content with more edges
for cname, gen in (("textured (a third of the texels carry an edge)", synth.textured_rgba8), ("uniform noise (84 %: nothing to skip)", synth.uniform_rgba8)):
    p = pool(gen, 1512, 1680, 8)
    report("C4 stereo 1512x1680->2016x2240 NIS r2.0, content: " + cname, dict(useNis=True, renderScale=0.75, sharpness=0.9, radius=2.0), p, 2 * 28224000, "pairs/s")
    del p
json.dump({"gpu": torch.cuda.get_device_name(0), "note": "1 GPU, device-resident, 2 frames in flight x 1 stream per eye; C4 on 2 GPUs (one eye each) is twice the per-GPU eye rate, C5 is C2 per GPU",
           "results": results}, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "configs_bench.json"), "w"), indent=1)
