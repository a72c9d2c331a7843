#!/usr/bin/env python
"""Build experimental variants of the library side by side (same sources, extra -D flags):
    python tools/build_variants.py name1:-DX=1,-DY=2 name2:-DZ=3 ...
-> vapor_b200/csrc/libvapor_b200_<name>.so, used with VAPOR_B200_LIB=<path>."""
import os, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vapor_b200 import _build

def one(spec):
    name, _, flags = spec.partition(":")
    out = os.path.join(_build.CSRC, f"libvapor_b200_{name}.so")
    _build.build_native(out=out, extra_flags=[f for f in flags.split(",") if f])
    return out

with ThreadPoolExecutor(4) as ex:
    for o in ex.map(one, sys.argv[1:]):
        print(o)
