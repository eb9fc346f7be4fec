"""include/ovrfsr.h is a C header: compile it alone as C11 and as C++17 with -Wall -Wextra -pedantic -Werror, and drive
the host-only entry points from a plain C program linked against the in-tree libovrfsr.so (no GPU involved)."""
import os
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
INC = ROOT / "include"


@pytest.mark.parametrize("compiler,std,lang", [("gcc", "-std=c11", "c"), ("g++", "-std=c++17", "c++")])
def test_header_compiles_alone(tmp_path, compiler, std, lang):
    if shutil.which(compiler) is None:
        pytest.skip(f"{compiler} not installed")
    src = tmp_path / ("h." + ("c" if lang == "c" else "cpp"))
    src.write_text('#include "ovrfsr.h"\nint main(void) { return (int)sizeof(ovrfsr_config) == 0; }\n')
    subprocess.check_call([compiler, std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(INC), "-c", str(src), "-o",
                           str(tmp_path / "h.o")])


def test_c_program_uses_the_abi(tmp_path, built_lib):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not installed")
    exe = tmp_path / "abi_smoke"
    libdir = Path(built_lib).parent
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(INC),
                           str(ROOT / "tests/c/abi_smoke.c"), "-o", str(exe), "-L", str(libdir), "-lovrfsr",
                           f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, env={**os.environ, "TZ": "UTC"})
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_smoke: ok" in out.stdout


def test_cpp_dropin_header_and_driver_compile(tmp_path):
    """The C++ drop-in header (vr::PostProcessor, Config singleton, the minimal OpenVR types) and the Submit-style
    driver of tests/cpp compile with a plain host compiler; the GPU test builds and runs the same file."""
    cuda_inc = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda")) / "include"
    if shutil.which("g++") is None or not (cuda_inc / "cuda_runtime.h").exists():
        pytest.skip("g++ or the CUDA headers are not installed")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", str(INC), "-I", str(ROOT / "openvr_fsr_b200/csrc"),
                           "-I", str(cuda_inc), "-c", str(ROOT / "tests/cpp/pp_selftest.cpp"), "-o", str(tmp_path / "pp.o")])
