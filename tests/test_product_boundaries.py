"""Structural guarantees of the product tree: nothing under vapor_b200/ imports or executes the oracle, the package
fails loudly without its CUDA library, and the C-ABI header and the ctypes binding list the same symbols."""
import ast
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vapor_b200")


def test_product_never_imports_the_oracle():
    for dirpath, _dirs, files in os.walk(PKG):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    assert not any(n == "oracle" or n.startswith("oracle.") for n in names), (path, names)
            elif f.endswith((".cu", ".cuh", ".h")):
                assert "oracle/" not in open(path).read(), path


def test_header_and_binding_agree():
    from vapor_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "vapor_b200.h")).read()
    declared = set(re.findall(r"\b(vapor_(?:gpu_\w+|hit_mix|host_plan|b200_abi_version))\s*\(", hdr))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vapor_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(_native.VaporNativeError):
        _native.load()


def test_no_gpu_means_no_results():
    """Without a CUDA device vapor_gpu_open fails and the engine raises: there is no CPU path to fall back to."""
    import pytest
    from vapor_b200 import _native
    from vapor_b200.engine import Engine
    lib = _native.load()
    if lib.vapor_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_native.VaporNativeError):
        Engine(0)
