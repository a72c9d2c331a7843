"""Compile the UNMODIFIED reference engine into oracle/_ref/ (git-ignored, travels with gpurun).

TEST INFRASTRUCTURE.  Recipe: Cython-compile /root/reference/vapor_vali/Simple_function.pyx *where it
lies* (language_level=3, exactly what the reference's own setup.py does with cythonize, setup.py:23)
into a C file under a temporary directory, then gcc it into oracle/_ref/vapor_ref_sf*.so.  Only the
compiled .so lands in the repo tree; no reference source is copied.  The module needs matplotlib at
import time (Simple_function.pyx:6-8); oracle/mpl_stub provides a no-op stand-in when it is absent.

Run:  python -m oracle.build_ref          (about 2 minutes, one core)
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/vapor_vali/Simple_function.pyx"
MODNAME = "vapor_ref_sf"


def built() -> bool:
    return bool(glob.glob(os.path.join(OUT, MODNAME + "*.so")))


def build(force: bool = False) -> str | None:
    if not os.path.exists(SRC):
        return None
    if built() and not force:
        return glob.glob(os.path.join(OUT, MODNAME + "*.so"))[0]
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="vapor_ref_build_")
    try:
        c_file = os.path.join(tmp, MODNAME + ".c")
        # cython names the module after the output file: -o <tmp>/vapor_ref_sf.c + --module-name
        subprocess.run([sys.executable, "-m", "cython", "-3", "--module-name", MODNAME, SRC, "-o", c_file],
                       check=True, cwd=tmp)
        ext = sysconfig.get_config_var("EXT_SUFFIX")
        so = os.path.join(OUT, MODNAME + ext)
        inc = sysconfig.get_paths()["include"]
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-w", "-I", inc, c_file, "-o", so], check=True)
        return so
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
