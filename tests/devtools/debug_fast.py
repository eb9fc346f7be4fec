"""Dev tool: where does fast math exceed 1 LSB per pass? (smoke case)"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
from oracle import pyoracle as po
dev = torch.device("cuda:0")
iw, ih, scale = 211, 157, 0.75
ow, oh = po.output_size(iw, ih, scale)
left, right = synth.stereo_pair("natural", iw, ih, 3)
for eye, img in ((0, left), (1, right)):
    for radius in (0.5, 2.0):
        uc = po.upscale_constants(eye, True, iw, ih, ow, oh, radius=radius)
        sc = po.sharpen_constants(eye, True, ow, oh, radius=radius, sharpness=0.9)
        easu = po.easu(img, ow, oh, uc); want = po.rcas(easu, sc)
        easu32 = po.easu(img, ow, oh, uc, out_dtype=np.float32); want32 = po.rcas(easu, sc, out_dtype=np.float32)
        t_src, t_easu = torch.from_numpy(img).to(dev), torch.from_numpy(easu).to(dev)
        for name, fn, src, ref8, ref32, shape in (("EASU", ovr.fsr_easu, t_src, easu, easu32, (oh, ow, 4)), ("RCAS", ovr.fsr_rcas, t_easu, want, want32, (oh, ow, 4))):
            consts = uc.words() if name == "EASU" else sc.words()
            g8 = torch.empty(shape, dtype=torch.uint8, device=dev); g32 = torch.empty(shape, dtype=torch.float32, device=dev)
            fn(src, g8, consts, ovr.MATH_FAST); fn(src, g32, consts, ovr.MATH_FAST); torch.cuda.synchronize()
            d8 = np.abs(g8.cpu().numpy().astype(np.int16) - ref8.astype(np.int16)); d32 = np.abs(g32.cpu().numpy() - ref32)
            print(f"eye {eye} r={radius} {name}: max8 {d8.max()} n>1 {(d8 > 1).sum()} max32 {np.nanmax(d32):.3e} nan {np.isnan(g32.cpu().numpy()).sum()}/{np.isnan(ref32).sum()}")
            ys, xs, cs = np.nonzero(d8 > 1)
            for y, x, c in list(zip(ys, xs, cs))[:5]:
                print("    at", x, y, c, "got", g8[y, x].tolist(), "ref", ref8[y, x].tolist(), "f32 got", g32[y, x].tolist(), "ref", ref32[y, x].tolist())
                if name == "RCAS":
                    print("      ring", easu[y-1, x, :3].tolist(), easu[y, x-1, :3].tolist(), easu[y, x, :3].tolist(), easu[y, x+1, :3].tolist(), easu[y+1, x, :3].tolist())
