"""Generates tests/golden/pass_*.npz from oracle/_ref (the reference's own NIS / CAS lines, and its FSR lines on
R10G10B10A2 textures) -- single-pass fixtures: input, constant block, reference output.

Run where /root/reference exists:   python tests/golden/make_golden_more.py
Checked by tests/test_golden_more.py (restated oracle, CPU) and tests/test_gpu_golden_more.py (CUDA path).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as po  # noqa: E402
from openvr_fsr_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def save(name, kind, src, out, consts, src_fmt=-1, dst_fmt=-1, sharpen_only=0):
    np.savez_compressed(OUT / f"pass_{name}.npz", kind=kind, src=src, out=out, consts=np.frombuffer(bytes(consts), dtype=np.uint32),
                        src_fmt=src_fmt, dst_fmt=dst_fmt, sharpen_only=sharpen_only)


def main():
    assert po.ref_available(), "needs /root/reference"
    # NIS NVScaler / NVSharpen (NIS_Scaler.h verbatim)
    for name, img, scale, radius, sharp, debug in (("nis_scaler_natural_61x45_s075", synth.natural_rgba8(61, 45, 2), 0.75, 2.0, 0.9, False),
                                                   ("nis_scaler_uniform_40x36_s059_r05_dbg", synth.uniform_rgba8(40, 36, 3), 0.59, 0.5, 0.4, True)):
        ih, iw = img.shape[:2]
        ow, oh = po.output_size(iw, ih, scale)
        cfg, _ = po.nis_config(False, 0, True, iw, ih, ow, oh, radius=radius, sharpness=sharp, debug=debug)
        save(name, "nis_scaler", img, po.nis_scaler(img, ow, oh, cfg, which="ref"), cfg)
    img = synth.natural_rgba8(70, 52, 4)
    cfg, _ = po.nis_config(True, 1, True, 70, 52, 70, 52, proj=(.45, .5, .55, .5), radius=0.6, sharpness=0.8)
    save("nis_sharpen_natural_70x52_r06", "nis_sharpen", img, po.nis_sharpen(img, cfg, which="ref"), cfg)
    # legacy CAS (ffx_cas.h CasFilter)
    img = synth.natural_rgba8(57, 41, 5)
    k = po.cas_setup(0.8, 1.0, 57, 41, 57, 41, which="ref")
    save("cas_sharpen_natural_57x41", "cas", img, po.cas(img, 57, 41, k, True, which="ref"), k, sharpen_only=1)
    k = po.cas_setup(1.0, 0.06, 57, 41, 57, 41, which="ref")
    save("cas_sharpen_clamped_57x41", "cas", img, po.cas(img, 57, 41, k, True, which="ref"), k, sharpen_only=1)
    ow, oh = po.output_size(57, 41, 0.67)
    k = po.cas_setup(0.9, 1.0, 57, 41, ow, oh, which="ref")
    save("cas_upscale_natural_57x41_s067", "cas", img, po.cas(img, ow, oh, k, False, which="ref"), k)
    f16 = synth.natural_rgba16f(33, 27, 6)
    ow, oh = po.output_size(33, 27, 0.5)
    k = po.cas_setup(0.5, 1.0, 33, 27, ow, oh, which="ref")
    save("cas_upscale_fp16_33x27_s05", "cas", f16, po.cas(f16, ow, oh, k, False, which="ref", out_dtype=np.float16), k)
    # FSR on R10G10B10A2 (DetermineOutputFormat keeps the 10-bit target)
    ten = synth.natural_rgb10a2(49, 37, 7)
    ow, oh = po.output_size(49, 37, 0.75)
    uc = po.upscale_constants(0, True, 49, 37, ow, oh, radius=0.5)
    e = po.easu(ten, ow, oh, uc, which="ref", src_fmt=po.FMT_RGB10A2, dst_fmt=po.FMT_RGB10A2)
    save("easu_rgb10a2_49x37_s075_r05", "easu", ten, e, uc, po.FMT_RGB10A2, po.FMT_RGB10A2)
    sc = po.sharpen_constants(0, True, ow, oh, radius=0.5, sharpness=0.9)
    save("rcas_rgb10a2_65x49_r05", "rcas", e, po.rcas(e, sc, which="ref", src_fmt=po.FMT_RGB10A2, dst_fmt=po.FMT_RGB10A2), sc,
         po.FMT_RGB10A2, po.FMT_RGB10A2)
    print("written:", sorted(p.name for p in OUT.glob("pass_*.npz")))


if __name__ == "__main__":
    main()
