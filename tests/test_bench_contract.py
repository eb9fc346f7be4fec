"""bench.py's command-line contract, as far as a box without a GPU can check it: the reference arm (the reference's own
lines on the host cores) prints ONE JSON line with the keys the driver reads, on the same config object as the product
arm; the product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    from oracle import pyoracle as po
    if not po.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("stereo eye-pairs/sec") and d["unit"] == "pairs/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and abs(d["value"] * d["ms_per_step"] * 1e-3 - 1.0) < 1e-6  # one pair per step
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("C2: stereo 1683x1869->2244x2492") and d["config"]["radius"] == 2.0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "pair" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0  # nothing of the product runs on this arm


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--steps", "1", "--warmup", "3", "--no-cpu-baseline", timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
    assert not [l for l in r.stdout.splitlines() if l.startswith('{"metric"')]
