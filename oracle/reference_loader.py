"""Import the UNMODIFIED reference module from /root/reference (build container only).

TEST INFRASTRUCTURE.  ``vapor_vali/Simple_function.pyx`` is untyped Python, so it imports as a plain
module (the survey confirmed the Cython build gives bit-identical results).  If a compiled copy exists
under ``oracle/_ref/`` (built by ``oracle/build_ref.py``) that one is preferred.  Nothing is copied into
the repo; on the GPU box ``/root/reference`` does not exist and ``load_reference`` returns None unless
the compiled ``oracle/_ref`` travelled with the snapshot.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/vapor_vali/Simple_function.pyx"
_cache = {}


def _ensure_stub():
    try:
        import matplotlib  # noqa: F401
    except Exception:
        stub = os.path.join(HERE, "mpl_stub")
        if stub not in sys.path:
            sys.path.insert(0, stub)


def load_reference(prefer_compiled: bool = True):
    """Return the reference ``Simple_function`` module, or None when it is not available here."""
    if "mod" in _cache:
        return _cache["mod"]
    _ensure_stub()
    mod = None
    ref_dir = os.path.join(HERE, "_ref")
    if prefer_compiled and os.path.isdir(ref_dir):
        sys.path.insert(0, ref_dir)
        try:
            mod = importlib.import_module("vapor_ref_sf")
            mod.__vapor_kind__ = "reference-cython"
        except Exception:
            mod = None
        finally:
            sys.path.remove(ref_dir)
    if mod is None and os.path.exists(REF_PYX):
        loader = importlib.machinery.SourceFileLoader("vapor_ref_sf_py", REF_PYX)
        spec = importlib.util.spec_from_loader("vapor_ref_sf_py", loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        mod.__vapor_kind__ = "reference-python"
    _cache["mod"] = mod
    return mod
