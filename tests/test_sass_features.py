"""Static check of the built library's machine code (cuobjdump, no GPU needed): the sm_100a features DESIGN.md names
are really in the shipped kernels, and the strict NIS kernels for UNORM sources carry no IEEE-division range check
(FCHK) -- a stale object file once made it look as if ptxas had turned div_rn_inrange back into div.rn."""
import re
import shutil
import subprocess

import pytest


@pytest.fixture(scope="module")
def sass(built_lib):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        r = subprocess.run([exe, "-sass", str(built_lib)], capture_output=True, text=True, timeout=600)
    except FileNotFoundError:
        pytest.skip("cuobjdump not installed")
    assert r.returncode == 0, r.stderr[-500:]
    funcs, name = {}, None
    for line in r.stdout.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and "/*" in line:
            funcs[name].append(line)
    return r.stdout, funcs


def _count(lines, op):
    return sum(1 for l in lines if re.search(r"\b" + re.escape(op), l))


def test_only_sm_100a_code_is_shipped(sass):
    text, _ = sass
    archs = set(re.findall(r"arch = (sm_\w+)", text))
    assert archs == {"sm_100a"}, archs


def test_tma_cluster_launch_control_and_packed_fp32_are_used(sass):
    _, funcs = sass
    def any_with(pattern, op):
        return [n for n, l in funcs.items() if re.search(pattern, n) and _count(l, op)]
    assert any_with(r"easu_kernelILi0ELi0ELi60ELb1", "UTMALDG")        # EASU, RGBA8, TMA variant
    assert any_with(r"rcas_kernelILi0ELi0ELb1", "UTMALDG")
    assert any_with(r"nis_scaler_kernelILi0ELi0ELb1", "UTMALDG")
    assert any_with(r"nis_scaler_kernel", "UGETNEXTWORKID")            # cluster launch control
    assert any_with(r"fast_math\d+easu_kernel", "FFMA2")               # packed FP32x2
    assert any_with(r"strict_math\d+nis_scaler_kernel", "FFMA2")
    assert any_with(r"fsr_fused_kernel", "STG.E.128")


def test_strict_nis_kernels_on_unorm_sources_have_no_division_range_check(sass):
    _, funcs = sass
    per = {}  # (kernel, source format) -> FCHK counts of its variants
    for name, lines in funcs.items():
        m = re.search(r"strict_math\d+(nis_scaler_kernel|nis_sharpen_kernel)ILi(\d+)E", name)
        if m:
            per.setdefault((m.group(1), int(m.group(2))), []).append(_count(lines, "FCHK"))
    unorm, flt = (0, 1, 4), (2, 3)  # RGBA8, BGRA8 (also BGRX8), RGB10A2 | RGBA16F, RGBA32F
    for f in unorm:
        assert max(per[("nis_sharpen_kernel", f)]) == 0
        # what is left in NVScaler are the outside-radius copy's coordinate quotients (x / radius.z, y / radius.w)
        assert max(per[("nis_scaler_kernel", f)]) <= 4
    for f in flt:  # float sources keep the checked IEEE divide: their operands are not range-bound
        assert min(per[("nis_sharpen_kernel", f)]) >= 4
        assert min(per[("nis_scaler_kernel", f)]) > max(per[("nis_scaler_kernel", 0)])
