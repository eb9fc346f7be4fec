"""In-tree build of the C-ABI library (openvr_fsr_b200/libovrfsr.so) for sm_100a.

nvcc cross-compiles without a GPU.  The kernels are compiled twice from the same sources:
kernels_fast.cu with -fmad=true and kernels_strict.cu with -fmad=false (see csrc/device_common.cuh).
cudart is linked statically and libcuda is only reached through cudaGetDriverEntryPoint at run
time, so the library loads (and exports every symbol of include/ovrfsr.h) on a machine without a driver.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libovrfsr.so"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]
UNITS = {
    "kernels_fast.cu": ["-fmad=true"],
    "kernels_strict.cu": ["-fmad=false"],
    "capi.cu": [],
    "nis_capi.cu": [],
    "postprocessor.cpp": [],
    "capture.cpp": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the sm_100a extension cannot be built")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc")) + \
        [PKG.parent / "include" / "ovrfsr.h", Path(__file__)]
    objs, jobs = [], []
    for name, extra in UNITS.items():
        src = CSRC / name
        if not src.exists():
            continue
        obj = objdir / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-x", "cu", "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            jobs.append(cmd)
        objs.append(obj)
    # the two kernel translation units dominate (~50 s each): compile the stale units side by side
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(4, os.cpu_count() or 1)) as pool:
        for rc in pool.map(lambda c: subprocess.call(c), jobs):
            if rc != 0:
                raise subprocess.CalledProcessError(rc, "nvcc")
    if force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-cudart", "static", "-o", str(LIB), *map(str, objs), "-lpthread", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
