"""NVScaler's two block schedules -- cluster launch control (default: one CTA per block, resident CTAs take over pending
ones) and the static round-robin walk (OVRFSR_NO_CLC=1, read once per process) -- must produce the same bits: the
schedule only decides WHICH CTA computes a block.  Each runs in its own interpreter because the switch is read once."""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

SCRIPT = r"""
import hashlib, sys
sys.path.insert(0, %r)
import numpy as np, torch
import openvr_fsr_b200 as ovr
from openvr_fsr_b200 import synth
dev = torch.device("cuda:0")
h = hashlib.sha256()
for (iw, ih, scale, radius) in ((611, 433, 0.75, 2.0), (611, 433, 0.75, 0.35), (300, 500, 0.59, 0.5)):
    ow, oh = ovr.output_size(iw, ih, scale)
    cfg = ovr.Config(fsrEnabled=True, useNis=True, renderScale=scale, sharpness=0.9, radius=radius, projCentre=(.45, .5, .55, .5))
    ncfg, _ = ovr.make_nis_config(cfg, False, 0, True, iw, ih, ow, oh)
    for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
        for src in (ovr.to_image(synth.natural_rgba8(iw, ih, 3), dev), torch.from_numpy(synth.natural_rgba8(iw, ih, 4)).to(dev)):
            out = torch.zeros((oh, ow, 4), dtype=torch.uint8, device=dev)
            for _ in range(2):
                ovr.nis_scaler(src, out, ncfg, mode)
            torch.cuda.synchronize()
            h.update(out.cpu().numpy().tobytes())
print(h.hexdigest())
""" % str(ROOT)


def _digest(env_extra):
    env = dict(os.environ, **env_extra)
    env.pop("OVRFSR_NO_CLC", None) if not env_extra else None
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()[-1]


def test_cluster_launch_control_and_static_schedule_agree(cuda):
    assert _digest({}) == _digest({"OVRFSR_NO_CLC": "1"})
