"""Build the CUDA C-ABI library in-tree: vapor_b200/csrc/libvapor_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libvapor_b200.so")
SOURCES = ["api.cu", "hostio.cpp"]
HEADERS = ["common.cuh", "k1_pack.cuh", "k2_tile.cuh", "k2_join.cuh", "k3_score.cuh", "k3_warp.cuh", "k4_genotype.cuh",
           os.path.join("..", "..", "include", "vapor_b200.h"), os.path.join("..", "..", "include", "vapor_hostio.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build vapor_b200/csrc/libvapor_b200.so")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_native(force: bool = False, verbose: bool = False, out: str = "", extra_flags=()) -> str:
    """Compile the kernels + C-ABI with nvcc for sm_100a.  Returns the library path.  ``out`` / ``extra_flags``
    build an experimental variant next to the product library (loaded with VAPOR_B200_LIB=...)."""
    if not out and not force and not is_stale():
        return LIB
    extra = os.environ.get("VAPOR_NVCC_EXTRA", "").split() + list(extra_flags)     # experiments only, e.g. -DK2_MINB=6
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", out or LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lz"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out or LIB


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
