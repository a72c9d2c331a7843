import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", params=["join", "tile"])
def engine(request):
    """One vapor_b200 handle on cuda:0 per kernel-2 variant for the whole GPU session (every GPU test runs against
    both the join kernel, the default, and the all-pairs tile kernel); fails loudly if the CUDA library is missing."""
    from vapor_b200.engine import Engine
    eng = Engine(0)
    eng.set_option("k2_mode", 1 if request.param == "join" else 0)
    eng.k2_mode_name = request.param
    yield eng
    eng.close()
