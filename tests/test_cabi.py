"""The C-ABI library loads and exports every symbol include/vapor_b200.h declares; without a GPU the
product path fails loudly instead of falling back to anything."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header="vapor_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vapor_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from vapor_b200 import _build, _native
    _build.build_native()
    lib = _native.load()
    declared = _declared_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vapor_b200.h but not exported"
    assert sorted(_native.EXPORTS) == declared
    assert lib.vapor_b200_abi_version() == 3
    # the native region extraction (include/vapor_hostio.h) lives in the same library
    from vapor_b200 import _hostio
    io_declared = _declared_functions("vapor_hostio.h")
    assert len(io_declared) >= 10
    for name in io_declared:
        assert hasattr(lib, name), f"{name} declared in include/vapor_hostio.h but not exported"
    assert sorted(_hostio.EXPORTS) == io_declared


def test_hit_mix_host_callable_matches_numpy():
    import numpy as np
    from vapor_b200 import _native
    from vapor_b200.engine import hit_mix
    lib = _native.load()
    rng = np.random.default_rng(0)
    xs = rng.integers(0, 2**27, 50); ys = rng.integers(0, 2**27, 50)
    for x, y in zip(xs, ys):
        assert lib.vapor_hit_mix(int(x), int(y)) == int(hit_mix(x, y))


def test_product_path_has_no_cpu_fallback():
    """No vapor_b200 module imports the oracle; opening a handle without a GPU raises."""
    import torch
    pkg = os.path.join(ROOT, "vapor_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("no oracle", ""), fn
    if not torch.cuda.is_available():
        from vapor_b200 import _native
        from vapor_b200.engine import Engine
        with pytest.raises(_native.VaporNativeError, match="no CPU fallback"):
            Engine(0)
