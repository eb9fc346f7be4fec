"""The CHECKER under AddressSanitizer + UndefinedBehaviorSanitizer: every pass of the restated oracle on small ragged
images (1x1 up to 64x24, every format, up- and down-scale factors) must run clean -- an out-of-bounds read in the
oracle would otherwise be able to 'agree' with anything.  Skipped where no compiler here ships the sanitizer runtimes."""
import shutil
import subprocess
from pathlib import Path

import pytest

ORACLE = Path(__file__).resolve().parents[1] / "oracle"


def test_oracle_runs_clean_under_asan_ubsan():
    built = False
    for cc in ("gcc", "/usr/bin/gcc", "cc", "clang"):
        path = shutil.which(cc) or (cc if Path(cc).exists() else None)
        if not path:
            continue
        r = subprocess.run(["make", "-C", str(ORACLE), "-B", "selfcheck", f"SANCC={path}"], capture_output=True, text=True)
        if r.returncode == 0:
            built = True
            break
    if not built:
        pytest.skip("no compiler with the ASan/UBSan runtimes available")
    out = subprocess.run([str(ORACLE / "selfcheck_asan")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "configurations ok" in out.stdout
