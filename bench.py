#!/usr/bin/env python
"""bench.py -- stereo eye-pairs/s of the FSR1 EASU+RCAS pass (BASELINE.json metric) on N B200s.

Workload (config.workload): BASELINE.json configs[1] = SURVEY.md C2: stereo 1683x1869 -> 2244x2492 RGBA8,
renderScale 0.75, sharpness 0.9, FSR path, radius 2.0 (mask off: EVERY pixel takes EASU+RCAS; the reference's
default radius 0.5 is reported beside it as `masked_r0.5`).  A "step" is one pass of the hot path over a batch
of POOL distinct stereo pairs per GPU; the input pool (POOL x 25 MB) is larger than the 126 MB L2, so no step
re-reads inputs from L2 (config.l2: "inputs larger than L2").

  value     : whole-job pairs/s, inputs resident in HBM, CUDA-event timed, max over ranks.  --streams (default 4):
              streams/2 PostProcessor contexts = frames in flight, one CUDA stream per eye of each; 1 = everything
              strictly back to back on one stream
  e2e       : same metric through PostProcessor.apply_host with pinned HOST buffers (H2D + kernels + D2H timed)
  roofline  : dominant kernel (EASU) algorithmic bytes / its mean launch time (CUDA events on the launch
              stream, second instrumented pass, the kernel alone on its stream) against MEASURED_PEAKS.json hbm_gbs;
              fp32_issue_frac = executed warp-instructions/s of that kernel against the SM issue peak
  issue_roofline_whole_step : executed warp-instructions per second of the whole step (ncu counts in
              profiles/kernel_constants.json x pairs/s) against 148 SMs x 4 schedulers x the sampled SM clock -- the
              bound that actually applies to the unmasked pass
  clocks    : nvidia-smi SM clock / throttle reasons sampled every 20 ms inside the timed region
  cpu_baseline : the reference's own lines (oracle/_ref, kind "reference") or the restated oracle ("port")
              on the box's host cores, one stereo pair, rank 0 at N=1 only
  --impl reference : the CPU reference arm (same metric/config), rank 0 only under torchrun

Multi-GPU: frames are independent -> each rank processes its own pool, no data-path
collective; NCCL only broadcasts the constant block from rank 0 and forms the barriers.  scaling = weak.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

IN_W, IN_H, RENDER_SCALE, SHARPNESS = 1683, 1869, 0.75, 0.9
OUT_W, OUT_H = 2244, 2492
EASU_BYTES_PER_EYE = IN_W * IN_H * 4 + OUT_W * OUT_H * 4          # 34,950,300
RCAS_BYTES_PER_EYE = 2 * OUT_W * OUT_H * 4                          # 44,736,384
PAIR_BYTES = 2 * (EASU_BYTES_PER_EYE + RCAS_BYTES_PER_EYE)          # 159,373,368
METRIC = "stereo eye-pairs/sec EASU+RCAS @2244x2492"
WORKLOAD = "C2: stereo 1683x1869->2244x2492 RGBA8, renderScale=0.75, sharpness=0.9, FSR EASU+RCAS"


def _profile_constants():
    """Per-launch DRAM traffic and executed instructions per output pixel of the shipped kernels on this workload,
    taken from the committed ncu captures (profiles/kernel_constants.json, written by tools/ncu_summary.py)."""
    p = ROOT / "profiles" / "kernel_constants.json"
    try:
        return json.loads(p.read_text())
    except Exception:
        return {}


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed workload runs (B200_PROFILING.md recipe).  The sampler
    is started before the warm-up (nvidia-smi needs ~100 ms to produce its first row); rows are time-stamped on
    arrival and only those inside the marked window [mark_begin, mark_end] count."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def samples_in_window(self):
        return sum(1 for ts, _ in self.rows if self.t0 is not None and ts >= self.t0 and (self.t1 is None or ts <= self.t1))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if self.t0 is None or ts < self.t0 or (self.t1 is not None and ts > self.t1):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# CPU reference arm / baseline
# ------------------------------------------------------------------------------------------------------
def cpu_pair_seconds(left, right, radius, which, threads, reps=2):
    """Time one stereo pair (EASU+RCAS, both eyes) on the host with `threads` threads; best of reps."""
    from oracle import pyoracle as po
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        for eye, img in ((0, left), (1, right)):
            uc = po.upscale_constants(eye, True, IN_W, IN_H, OUT_W, OUT_H, radius=radius)
            sc = po.sharpen_constants(eye, True, OUT_W, OUT_H, radius=radius, sharpness=SHARPNESS)
            po.rcas(po.easu(img, OUT_W, OUT_H, uc, which=which, nthreads=threads), sc, which=which, nthreads=threads)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_baseline(radius, reps=2):
    from oracle import pyoracle as po
    from openvr_fsr_b200 import synth
    which, kind = ("ref", "reference") if po.ref_available() else ("oracle", "port")
    threads = os.cpu_count() or 1
    left, right = synth.stereo_pair("natural", IN_W, IN_H, 1)
    sec = cpu_pair_seconds(left, right, radius, which, threads, reps)
    return {"value": 1.0 / sec, "unit": "pairs/s", "cores": threads, "kind": kind,
            "sample": f"1 stereo pair of the same workload (both eyes EASU+RCAS), {threads} threads, best of {reps}"}, which


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    from openvr_fsr_b200 import synth
    which, kind = ("ref", "reference") if po.ref_available() else ("oracle", "port")
    threads = os.cpu_count() or 1
    left, right = synth.stereo_pair("natural", IN_W, IN_H, 1)
    for _ in range(min(args.warmup, 1)):
        cpu_pair_seconds(left, right, args.radius, which, threads, 1)
    steps = max(1, min(args.steps, 12))  # each step = 1 pair; bounded so the run ends within minutes
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_pair_seconds(left, right, args.radius, which, threads, 1)
    dt = time.perf_counter() - t0
    val = steps / dt
    sample = f"{steps} steps x 1 stereo pair, {threads} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "radius": args.radius, "pairs_per_step": 1},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--radius", type=float, default=2.0, help="Config::radius; 2.0 = mask off (headline)")
    ap.add_argument("--pool", type=int, default=8, help="distinct stereo pairs per GPU per step (pool > L2)")
    ap.add_argument("--math", default="strict", choices=["fast", "strict"],
                    help="strict = bit-identical to the reference lines end to end (headline); fast = <=1 LSB per pass")
    ap.add_argument("--streams", type=int, default=4, choices=[1, 2, 4, 6, 8],
                    help="1 = everything on one stream; 2 = one stream per eye; 4+ = streams/2 frames in flight, one context each")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    math_mode = ovr.MATH_STRICT if args.math == "strict" else ovr.MATH_FAST
    cfg = ovr.Config(fsrEnabled=True, renderScale=RENDER_SCALE, sharpness=SHARPNESS, radius=args.radius,
                     mathMode=math_mode, device=local_rank)

    # north_star: NCCL only as barrier / broadcast of the shared FSR constants (96 + 48 bytes), root 0
    from openvr_fsr_b200 import sharding
    consts = sharding.broadcast_constants(cfg, IN_W, IN_H, OUT_W, OUT_H, dev if world > 1 else None)

    # per-rank input pool: POOL distinct stereo pairs, resident in HBM
    pool = []
    base_l, base_r = synth.stereo_pair("natural", IN_W, IN_H, 1)
    for i in range(args.pool):
        sh = 37 * (i + rank * args.pool)
        # device images with a 256-byte-aligned row pitch (what cudaMallocPitch / ovrfsr_image_alloc return): the
        # TMA tile loader needs a 16-byte-aligned pitch; algorithmic bytes are counted without the padding
        pool.append((ovr.to_image(np.roll(base_l, sh, axis=0), dev), ovr.to_image(np.roll(base_r, sh, axis=0), dev)))
    runner = EyeStreams(ovr, torch, cfg, dev, args.streams)
    pp = runner.pps[0]
    assert np.array_equal(pp_consts_after_first(pp, pool[0][0]), consts["upscale"][0]), "rank constants differ from root's"
    step, fork, join = (lambda: runner.step(pool)), runner.fork, runner.join

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.mark_begin()
    launches0 = ovr.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    fork()
    for _ in range(args.steps):
        step()
    join()
    ev1.record()
    barrier()
    launches = ovr.kernel_launches() - launches0
    clock_note = "sampled inside the timed region"
    if rank == 0 and sampler.samples_in_window() < 5:
        # a short timed region (tens of ms) can fall between two nvidia-smi rows: keep the SAME steps running,
        # untimed, until a handful of rows under load exist
        clock_note = "timed region shorter than the sampling period: sampled during identical untimed steps run right after it"
        t_end = time.perf_counter() + 1.5
        while sampler.samples_in_window() < 5 and time.perf_counter() < t_end:
            step()
            torch.cuda.synchronize()
    sampler.mark_end()
    elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed_ms.item())
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["note"] = clock_note
    pairs = world * args.pool * args.steps
    value = pairs / (elapsed_ms * 1e-3)

    # ---- instrumented pass: per-kernel CUDA events on the launch stream (roofline of the dominant kernel)
    easu_ms, rcas_ms = per_kernel_times(ovr, pool, consts, math_mode, min(args.steps, 5))
    peak, peak_src = _peaks()
    easu_gbs = EASU_BYTES_PER_EYE / (easu_ms * 1e-3) / 1e9
    rcas_gbs = RCAS_BYTES_PER_EYE / (rcas_ms * 1e-3) / 1e9
    prof_all = _profile_constants()
    prof = {k[: -len(args.math) - 1]: v for k, v in prof_all.items() if k.endswith("_" + args.math)}
    sm_clock_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    def issue_frac(instr_per_px, ms):  # executed warp-instructions / s against 4 issue slots / SM / clock
        if not instr_per_px:
            return None
        return instr_per_px * OUT_W * OUT_H / 32 / (ms * 1e-3) / (148 * 4 * sm_clock_mhz * 1e6)
    roofline = {"bound": "hbm", "kernel": "easu_kernel", "achieved": easu_gbs, "peak": peak, "unit": "GB/s",
                "frac": easu_gbs / peak, "traffic": prof.get("easu_traffic_bytes"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": EASU_BYTES_PER_EYE, "ms_per_launch": easu_ms,
                "fp32_issue_frac": issue_frac(prof.get("easu_instr_per_px"), easu_ms),
                "instr_per_output_px": prof.get("easu_instr_per_px"),
                "timing": "ms_per_launch = this kernel alone on its stream (CUDA events); inside the step, kernels of "
                          "the other eye / frame run concurrently, so ms_per_step is below the sum of launch times",
                "note": "EASU is FP32-issue-bound, not HBM-bound, when the mask is off (DESIGN.md section 5): "
                        "fp32_issue_frac = executed warp-instructions/s (ncu count x live launch rate) / issue peak"}
    roofline_rcas = {"bound": "hbm", "kernel": "rcas_kernel", "achieved": rcas_gbs, "peak": peak, "unit": "GB/s",
                     "frac": rcas_gbs / peak, "traffic": prof.get("rcas_traffic_bytes"),
                     "algorithmic_bytes_per_launch": RCAS_BYTES_PER_EYE, "ms_per_launch": rcas_ms,
                     "fp32_issue_frac": issue_frac(prof.get("rcas_instr_per_px"), rcas_ms),
                     "instr_per_output_px": prof.get("rcas_instr_per_px")}

    # ---- end to end: host buffers through the public API, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = e2e_run(ovr, torch, dist, cfg, pool, dev, world, max(2, args.steps // 4), args.warmup)

    # ---- reference default radius beside the headline, and the other math mode
    masked = None
    if args.radius != 0.5:
        masked = quick_value(ovr, torch, cfg, pool, max(3, args.steps // 2), args.streams, radius=0.5)
    other_mode = "fast" if args.math == "strict" else "strict"
    other = quick_value(ovr, torch, cfg, pool, max(3, args.steps // 2), args.streams,
                        mathMode=ovr.MATH_FAST if other_mode == "fast" else ovr.MATH_STRICT)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_baseline(args.radius)

    runner.close()
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "radius": args.radius, "pairs_per_step_per_gpu": args.pool,
                       "math": args.math, "source_pitch": "256-byte aligned rows (TMA tile loads)", "l2": f"inputs larger than L2 ({args.pool} distinct pairs = "
                       f"{args.pool * 2 * IN_W * IN_H * 4 / 1e6:.0f} MB per GPU per step)",
                       "streams": f"{args.streams} CUDA streams per GPU: {max(1, args.streams // 2)} frame(s) in flight, one stream per eye",
                       "parallelism": f"frames sharded {world}x, no data-path collective"},
            "hbm_gbs_whole_pass": value / world * PAIR_BYTES / 1e9,
            "hbm_frac_whole_pass": value / world * PAIR_BYTES / 1e9 / peak,
            "roofline": roofline, "roofline_rcas": roofline_rcas, "clocks": clocks, "gpu_launches": int(launches),
        }
        if prof.get("easu_instr_per_px") and prof.get("rcas_instr_per_px"):
            # the bound that actually applies to the unmasked pass: warp-instruction issue slots (4 per SM per clock)
            wi = (prof["easu_instr_per_px"] + prof["rcas_instr_per_px"]) * OUT_W * OUT_H * 2 / 32 * (value / world)
            out["issue_roofline_whole_step"] = {
                "achieved": wi / 1e12, "peak": 148 * 4 * sm_clock_mhz * 1e6 / 1e12, "unit": "T warp-instr/s",
                "frac": wi / (148 * 4 * sm_clock_mhz * 1e6),
                "note": "executed warp-instructions per pair (ncu counts in profiles/kernel_constants.json) x pairs/s, "
                        "against 148 SMs x 4 schedulers x the SM clock sampled during the run"}
        if e2e is not None:
            out["e2e"] = e2e
        if masked is not None:
            out["masked_r0.5"] = {"value": masked * world, "unit": "pairs/s",
                                  "note": "reference default radius 0.5 (EASU/RCAS inside the radius only)"}
        out[f"value_{other_mode}_math"] = {"value": other * world, "unit": "pairs/s", "note": (
            "FMA-contracted kernels: each pass <= 1 LSB from the reference lines on identical inputs" if other_mode == "fast"
            else "reference operation order: bit-identical to the reference lines end to end")}
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def pp_consts_after_first(pp, tex):
    import torch
    pp.apply(0, tex)
    torch.cuda.synchronize()
    c = pp.upscale_constants(0)
    pp.reset()
    return c


def per_kernel_times(ovr, pool, consts, math_mode, reps):
    """Mean device time of one EASU and one RCAS launch (per eye), events recorded on the launching stream."""
    import torch
    dev = pool[0][0].device
    mid = ovr.alloc_image(OUT_W, OUT_H, torch.uint8, dev)
    dst = ovr.alloc_image(OUT_W, OUT_H, torch.uint8, dev)
    marks = []  # (e0, e1, e2) per eye; nothing synchronises inside the loop, so the GPU stays busy
    for _ in range(reps):
        for left, right in pool:
            for eye, tex in ((0, left), (1, right)):
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                evs[0].record()
                ovr.fsr_easu(tex, mid, consts["upscale"][eye], math_mode)
                evs[1].record()
                ovr.fsr_rcas(mid, dst, consts["sharpen"][eye], math_mode)
                evs[2].record()
                marks.append(evs)
    torch.cuda.synchronize()
    te = [m[0].elapsed_time(m[1]) for m in marks]
    tr = [m[1].elapsed_time(m[2]) for m in marks]
    skip = min(len(te) // 4, 8)
    return statistics.mean(te[skip:]), statistics.mean(tr[skip:])


class EyeStreams:
    """The device-resident workload driver: n_streams // 2 PostProcessor contexts (frames in flight), one CUDA stream
    per eye of each.  A context keeps one output set per eye, so its two eyes can run concurrently; frame i goes to
    context i % n_ctx.  Kernels of different eyes / frames then share the SMs: a persistent EASU grid's last wave no
    longer leaves SMs idle, and issue-bound EASU warps interleave with the other eye's RCAS warps."""

    def __init__(self, ovr, torch, cfg, dev, n_streams):
        self.main = torch.cuda.current_stream(dev)
        self.pps = [ovr.PostProcessor(cfg) for _ in range(max(1, n_streams // 2))]
        if n_streams == 1:
            self.streams = [[self.main, self.main]]
        else:
            self.streams = [[torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)] for _ in self.pps]

    def step(self, pool):
        n = len(self.pps)
        for i, (left, right) in enumerate(pool):
            self.pps[i % n].apply(0, left, stream=self.streams[i % n][0])
            self.pps[i % n].apply(1, right, stream=self.streams[i % n][1])

    def fork(self):
        for pair in self.streams:
            for s in pair:
                if s is not self.main:
                    s.wait_stream(self.main)

    def join(self):
        for pair in self.streams:
            for s in pair:
                if s is not self.main:
                    self.main.wait_stream(s)

    def close(self):
        for p in self.pps:
            p.close()


def quick_value(ovr, torch, cfg, pool, steps, n_streams=2, **changes):
    import dataclasses
    dev = pool[0][0].device
    r = EyeStreams(ovr, torch, dataclasses.replace(cfg, **changes), dev, n_streams)
    for _ in range(2):
        r.step(pool)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r.fork()
    for _ in range(steps):
        r.step(pool)
    r.join()
    e1.record()
    torch.cuda.synchronize()
    v = len(pool) * steps / (e0.elapsed_time(e1) * 1e-3)
    r.close()
    return v


def e2e_run(ovr, torch, dist, cfg, pool, dev, world, steps, warmup):
    """Same metric through the reference-facing call with HOST buffers: per eye, pinned host -> device copy,
    EASU+RCAS, device -> pinned host copy, all inside the timed region.  Two PostProcessor contexts are ping-ponged
    (frame i uses context i % 2, each with one stream per eye) so that the upload of frame i+1 overlaps the download
    of frame i: a context owns its staging / output images, so frames in flight need one context each."""
    pps = [ovr.PostProcessor(cfg), ovr.PostProcessor(cfg)]
    n = len(pool)
    h_in = [(l.cpu().contiguous().pin_memory(), r.cpu().contiguous().pin_memory()) for l, r in pool]
    h_out = [(torch.empty((OUT_H, OUT_W, 4), dtype=torch.uint8).pin_memory(),
              torch.empty((OUT_H, OUT_W, 4), dtype=torch.uint8).pin_memory()) for _ in range(n)]
    streams = [[torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)] for _ in pps]

    def step():
        for i in range(n):
            for eye in (0, 1):
                pps[i & 1].apply_host(eye, h_in[i][eye], h_out[i][eye], stream=streams[i & 1][eye])

    for _ in range(max(1, min(warmup, 2))):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0][0])
    for _ in range(steps):
        step()
    for s in (streams[0][1], streams[1][0], streams[1][1]):
        streams[0][0].wait_stream(s)
    e1.record(streams[0][0])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    for p in pps:
        p.close()
    pairs = world * n * steps
    return {"value": pairs / (float(ms.item()) * 1e-3), "unit": "pairs/s",
            "h2d_bytes_per_step": n * 2 * IN_W * IN_H * 4, "d2h_bytes_per_step": n * 2 * OUT_W * OUT_H * 4,
            "steps": steps, "wall_value": pairs / world / wall * world,
            "note": "pinned host -> H2D -> EASU+RCAS -> D2H per eye via ovrfsr_apply_host; 2 contexts x 2 streams in flight"}


if __name__ == "__main__":
    main()
