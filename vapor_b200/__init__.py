"""vapor_b200: B200-native drop-in for VaPoR's per-read scoring hot path.

The package holds only what that path needs: ``csrc/`` (hand-written sm_100a CUDA
kernels behind the C-ABI of ``include/vapor_b200.h``), the ctypes binding
(``_native``), the batching engine (``engine``) and the host-side mirror of the
reference interface (``Simple_function``, ``prep``, ``cli``).  There is no CPU
implementation of the scoring path in here.
"""
__version__ = "0.1.0"
