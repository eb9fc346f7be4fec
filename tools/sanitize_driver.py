"""Driver for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): one launch of every kernel variant on
small images with ragged edges -- TMA and plain loaders, both math modes, every format, masked and unmasked, NIS, the
MSAA resolve and the host entry.  Run as:  compute-sanitizer --tool racecheck python tools/sanitize_driver.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth

dev = torch.device("cuda:0")
n = 0
for math in (ovr.MATH_STRICT, ovr.MATH_FAST):
    for (iw, ih, scale) in ((150, 101, 0.75), (97, 64, 0.59), (80, 70, 1.0), (66, 50, 1.3)):
        ow, oh = ovr.output_size(iw, ih, scale)
        for radius in (2.0, 0.4):
            cfg = ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=radius, mathMode=math)
            uc = ovr.make_upscale_constants(cfg, 0, True, iw, ih, ow, oh)
            sc = ovr.make_sharpen_constants(cfg, 0, True, ow, oh)
            srcs = [(ovr.to_image(synth.natural_rgba8(iw, ih, 1), dev), None, "rgba8 aligned (TMA)"),
                    (torch.from_numpy(synth.natural_rgba8(iw, ih, 2)).to(dev), None, "rgba8 tight (plain loads)"),
                    (ovr.to_image(synth.natural_rgba8(iw, ih, 3), dev), ovr.FORMAT_BGRA8, "bgra8"),
                    (ovr.to_image(synth.natural_rgb10a2(iw, ih, 4), dev), ovr.FORMAT_RGB10A2, "rgb10a2"),
                    (ovr.to_image(synth.natural_rgba16f(iw, ih, 5), dev), None, "rgba16f")]
            for src, fmt, _ in srcs:
                ten = fmt == ovr.FORMAT_RGB10A2
                mid = ovr.alloc_image(ow, oh, torch.uint8, dev)
                out = ovr.alloc_image(ow, oh, torch.uint8, dev)
                if scale != 1.0 and src.dtype == torch.uint8:  # round 2: the fused EASU->RCAS kernel on the same inputs
                    fz = ovr.alloc_image(ow, oh, torch.uint8, dev)
                    ovr.fsr_fused(src, fz, uc, sc, math, src_fmt=fmt, dst_fmt=ovr.FORMAT_RGB10A2 if ten else None)
                    n += 1
                if scale != 1.0:
                    ovr.fsr_easu(src, mid, uc, math, src_fmt=fmt, dst_fmt=ovr.FORMAT_RGB10A2 if ten else None)
                    ovr.fsr_rcas(mid, out, sc, math, src_fmt=ovr.FORMAT_RGB10A2 if ten else None,
                                 dst_fmt=ovr.FORMAT_RGB10A2 if ten else None)
                    n += 2
                else:
                    ovr.fsr_rcas(src, out, sc, math, src_fmt=fmt, dst_fmt=ovr.FORMAT_RGB10A2 if ten else None)
                    n += 1
                if fmt is None and src.dtype == torch.uint8:
                    f16 = ovr.alloc_image(ow, oh, torch.float16, dev)
                    ovr.fsr_rcas(out, f16, sc, math)
                    n += 1
            # legacy CAS
            if scale <= 1.0:
                src = ovr.to_image(synth.natural_rgba8(iw, ih, 9), dev)
                out = ovr.alloc_image(ow, oh, torch.uint8, dev)
                if scale == 1.0:
                    ovr.cas(src, out, ovr.cas_setup(0.8, 1.0, iw, ih, iw, ih), True, math)
                    ovr.cas(torch.from_numpy(synth.natural_rgba8(iw, ih, 10)).to(dev), out, ovr.cas_setup(0.8, 0.5, iw, ih, iw, ih), True, math)
                    n += 2
                else:
                    ovr.cas(src, out, ovr.cas_setup(0.8, 1.0, iw, ih, ow, oh), False, math)
                    n += 1
            # NIS
            if scale <= 1.0:
                ncfg, _ = ovr.make_nis_config(ovr.Config(fsrEnabled=True, useNis=True, renderScale=scale, sharpness=0.7, radius=radius),
                                              scale == 1.0, 0, True, iw, ih, ow, oh)
                src = ovr.to_image(synth.natural_rgba8(iw, ih, 6), dev)
                out = ovr.alloc_image(ow, oh, torch.uint8, dev)
                (ovr.nis_sharpen if scale == 1.0 else ovr.nis_scaler)(src, out, ncfg, math)
                n += 1
    torch.cuda.synchronize()
# round 2: paired passes through the context (masked and unmasked), BGRX8 / RGB32F sources, NVScaler over many blocks
for math in (ovr.MATH_STRICT, ovr.MATH_FAST):
    for radius in (2.0, 0.35):
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=radius, mathMode=math))
        pp.apply(0, ovr.to_image(synth.natural_rgba8(331, 203, 11), dev))
        pp.apply(1, ovr.to_image(synth.natural_rgba8(331, 203, 12), dev), fmt=ovr.FORMAT_BGRX8 | ovr.FORMAT_SRGB_BIT)
        pp.close()
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=radius, mathMode=math, fusedFsr=True))
        pp.apply(0, ovr.to_image(synth.natural_rgba8(331, 203, 13), dev))
        pp.close()
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, useNis=True, renderScale=0.75, sharpness=0.9, radius=radius, mathMode=math))
        pp.apply(0, ovr.to_image(synth.natural_rgba8(331, 203, 14), dev))
        pp.apply(1, torch.from_numpy(synth.natural_rgba8(331, 203, 15)).to(dev), fmt=ovr.FORMAT_BGRX8)
        pp.apply(0, torch.from_numpy(np.ascontiguousarray(synth.natural_rgba16f(120, 90, 2).astype(np.float32)[..., :3])).to(dev), fmt=ovr.FORMAT_RGB32F)
        pp.close()
    torch.cuda.synchronize()
# round 2 (later): both eyes per call on two streams (ovrfsr_apply_pair), NIS on edge-free / mixed / edge-dense content
for kw in (dict(renderScale=0.75, radius=0.5), dict(renderScale=0.75, radius=0.5, useNis=True), dict(renderScale=1.0, radius=2.0, useNis=True)):
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, sharpness=0.9, **kw))
    for i in range(3):
        pp.apply_pair(ovr.to_image(synth.textured_rgba8(331, 203, 20 + i, cell=16), dev), ovr.to_image(synth.natural_rgba8(331, 203, 30 + i), dev))
    torch.cuda.synchronize()
    pp.close()
for gen in (synth.textured_rgba8, synth.uniform_rgba8, lambda w, h, s: np.full((h, w, 4), 90, np.uint8)):
    for scale in (0.75, 1.0):
        ow, oh = ovr.output_size(203, 131, scale)
        ncfg, _ = ovr.make_nis_config(ovr.Config(fsrEnabled=True, useNis=True, renderScale=scale, sharpness=0.7, radius=2.0), scale == 1.0, 0, True, 203, 131, ow, oh)
        out = ovr.alloc_image(ow, oh, torch.uint8, dev)
        (ovr.nis_sharpen if scale == 1.0 else ovr.nis_scaler)(ovr.to_image(gen(203, 131, 5), dev), out, ncfg, ovr.MATH_STRICT)
        n += 1
torch.cuda.synchronize()
# front end + stateful path + host entry
ms = torch.from_numpy(np.repeat(synth.natural_rgba8(90, 61, 7), 4, axis=1)).to(dev)
pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5))
pp.apply(0, ms, samples=4)
pp.apply(1, ms, samples=4)
h_in = torch.from_numpy(synth.natural_rgba8(90, 61, 8)).pin_memory()
ow, oh = ovr.output_size(90, 61, 0.75)
h_out = torch.empty((oh, ow, 4), dtype=torch.uint8).pin_memory()
pp.apply_host(0, h_in, h_out)
torch.cuda.synchronize()
pp.close()
print("sanitize_driver: %d stateless launches + ctx path done, library launches = %d" % (n, ovr.kernel_launches()))
