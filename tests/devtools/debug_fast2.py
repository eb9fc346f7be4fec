import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
from oracle import pyoracle as po
dev = torch.device("cuda:0")
iw, ih, scale = 200, 150, 0.77
ow, oh = po.output_size(iw, ih, scale)
for radius in (2.0, 0.5):
    uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=radius)
    sc = po.sharpen_constants(0, True, ow, oh, radius=radius, sharpness=0.9)
    for name, src in (("uniform", synth.uniform_rgba8(iw, ih, 0)), ("natural", synth.natural_rgba8(iw, ih, 1))):
        easu = po.easu(src, ow, oh, uc); want = po.rcas(easu, sc)
        easu32 = po.easu(src, ow, oh, uc, out_dtype=np.float32); want32 = po.rcas(easu, sc, out_dtype=np.float32)
        for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
            for pname, fn, s_, ref8, ref32, consts in (("EASU", ovr.fsr_easu, src, easu, easu32, uc.words()), ("RCAS", ovr.fsr_rcas, easu, want, want32, sc.words())):
                t = torch.from_numpy(s_).to(dev)
                g8 = torch.empty((oh, ow, 4), dtype=torch.uint8, device=dev); g32 = torch.empty((oh, ow, 4), dtype=torch.float32, device=dev)
                fn(t, g8, consts, mode); fn(t, g32, consts, mode); torch.cuda.synchronize()
                d8 = np.abs(g8.cpu().numpy().astype(np.int16) - ref8.astype(np.int16)); d32 = np.abs(g32.cpu().numpy() - ref32)
                print(f"r={radius} {name} mode={mode} {pname}: max8 {d8.max()} n>0 {(d8>0).sum()} n>1 {(d8>1).sum()} max32 {np.nanmax(d32):.3e}")
                ys, xs, cs = np.nonzero(d8 > 1)
                for y, x, c in list(zip(ys, xs, cs))[:3]:
                    print("    at", x, y, c, "got", g8[y, x].tolist(), "ref", ref8[y, x].tolist(), "f32", g32[y, x].tolist(), ref32[y, x].tolist())
