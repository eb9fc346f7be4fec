"""Generates tests/golden/*.npz from oracle/_ref (the reference's own kernel lines compiled on the host).

Run in a container where /root/reference exists:   python tests/golden/make_golden.py
Each fixture stores the input, the constant words and the reference output, so the fixtures pin BOTH the
restated oracle (tests/test_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py) on machines
where the reference is absent.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as po  # noqa: E402
from openvr_fsr_b200 import synth  # noqa: E402
from tests.cases import corner_images  # noqa: E402

OUT = Path(__file__).resolve().parent


def fsr_fixture(name, src, scale, radius, sharpness, eye=0, proj=(.5, .5, .5, .5), debug=False, src_fmt=None,
                out_dtype=np.uint8):
    ih, iw = src.shape[:2]
    ow, oh = po.output_size(iw, ih, scale)
    uc = po.upscale_constants(eye, True, iw, ih, ow, oh, proj=proj, radius=radius)
    sc = po.sharpen_constants(eye, True, ow, oh, proj=proj, radius=radius, sharpness=sharpness, debug=debug)
    easu = po.easu(src, ow, oh, uc, which="ref", src_fmt=src_fmt, out_dtype=out_dtype)
    rcas = po.rcas(easu, sc, which="ref", out_dtype=out_dtype)
    np.savez_compressed(OUT / f"fsr_{name}.npz", src=src, src_fmt=-1 if src_fmt is None else src_fmt,
                        upscale=uc.words(), sharpen=sc.words(), easu=easu, rcas=rcas)


def main():
    assert po.ref_available(), "needs /root/reference"
    imgs = corner_images(33, 47)
    fsr_fixture("natural_33x47_s075_r20", imgs["natural"], 0.75, 2.0, 0.9)
    fsr_fixture("uniform_33x47_s075_r05", imgs["uniform"], 0.75, 0.5, 0.9)
    fsr_fixture("checker_33x47_s075_r20", imgs["checker"], 0.75, 2.0, 1.0)
    fsr_fixture("diag_17x13_s05_r20", corner_images(17, 13)["diag"], 0.5, 2.0, 0.75)
    fsr_fixture("impulse_16x16_s05_r20", corner_images(16, 16)["impulse"], 0.5, 2.0, 0.9)
    fsr_fixture("const0_17x13_s075_r20", corner_images(17, 13)["const0"], 0.75, 2.0, 0.9)
    fsr_fixture("natural_96x80_s077_r04_dbg", synth.natural_rgba8(96, 80, 9), 0.77, 0.4, 0.9, eye=1,
                proj=(.45, .52, .55, .48), debug=True)
    fsr_fixture("bgra_40x24_s13_r20", synth.natural_rgba8(40, 24, 4), 1.3, 2.0, 0.5, src_fmt=po.FMT_BGRA8)
    fsr_fixture("fp16_40x24_s075_r20_f16out", synth.natural_rgba16f(40, 24, 6), 0.75, 2.0, 0.9, out_dtype=np.float16)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
