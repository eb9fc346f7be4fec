"""The C++ drop-in (vr::PostProcessor::Apply / Reset + Config singleton, openvr_fsr_b200/csrc/postprocessor.h) driven
the way VrHooks.cpp drives the reference class; its output must be bit-identical to the oracle (strictMath)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def selftest_exe(built_lib, tmp_path_factory):
    exe = tmp_path_factory.mktemp("cpp") / "pp_selftest"
    pkg = ROOT / "openvr_fsr_b200"
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-I", str(pkg / "csrc"), "-I", str(ROOT / "include"),
                           str(ROOT / "tests/cpp/pp_selftest.cpp"), "-o", str(exe), "-L", str(pkg), "-lovrfsr",
                           "-Xlinker", f"-rpath={pkg}"])
    return exe


@pytest.mark.parametrize("use_nis,scale", [(0, 0.75), (1, 0.75), (1, 1.0), (0, 1.0)])
def test_cpp_postprocessor_matches_oracle(cuda, selftest_exe, tmp_path, use_nis, scale):
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih, sharp, radius = 180, 132, 0.9, 0.5
    left, right = synth.stereo_pair("natural", iw, ih, 11)
    (tmp_path / "in.rgba").write_bytes(left.tobytes())
    outs = [tmp_path / "l.rgba", tmp_path / "r.rgba"]
    capdir = tmp_path / "captures"
    capdir.mkdir()
    subprocess.check_call([str(selftest_exe), str(tmp_path / "in.rgba"), str(iw), str(ih), str(scale), str(sharp),
                           str(radius), str(use_nis), str(outs[0]), str(outs[1]), str(capdir)])
    ow, oh = po.output_size(iw, ih, scale)
    for eye, img in ((0, left), (1, right)):
        got = np.frombuffer(outs[eye].read_bytes(), dtype=np.uint8).reshape(oh, ow, 4)
        if use_nis:
            cfg, _ = po.nis_config(scale == 1.0, eye, True, iw, ih, ow, oh, radius=radius, sharpness=sharp)
            want = po.nis_sharpen(img, cfg) if scale == 1.0 else po.nis_scaler(img, ow, oh, cfg)
        else:
            sc = po.sharpen_constants(eye, True, ow, oh, radius=radius, sharpness=sharp)
            mid = img if scale == 1.0 else po.easu(img, ow, oh, po.upscale_constants(eye, True, iw, ih, ow, oh, radius=radius))
            want = po.rcas(mid, sc)
        assert np.array_equal(got, want)
        if eye == 0:  # TakeCapture(): the F7 file holds the left eye's output (PostProcessor.cpp:634-657)
            import openvr_fsr_b200 as ovr
            files = list(capdir.iterdir())
            assert len(files) == 1 and files[0].name.endswith("_%s_s90_r50.dds" % ("nis" if use_nis else "fsr"))
            cap, fmt = ovr.load_dds(files[0])
            assert fmt == ovr.FORMAT_RGBA8 and np.array_equal(cap, want)
