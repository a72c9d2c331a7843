"""ctypes binding of the C-ABI in ``include/vapor_b200.h``.

Loading fails loudly (``VaporNativeError``) when the CUDA library has not been built or
cannot be loaded; there is no Python/CPU substitute for it in this package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VAPOR_B200_LIB: another in-tree build of the same sources (kernel experiments: tools/build_variants.sh)
LIB_PATH = os.environ.get("VAPOR_B200_LIB") or os.path.join(HERE, "csrc", "libvapor_b200.so")

VAPOR_MODE_ABS, VAPOR_MODE_W10, VAPOR_MODE_REDEF, VAPOR_MODE_ABS_AND_W10 = 0, 1, 2, 3
VAPOR_ST_SKIPPED, VAPOR_ST_SCORED, VAPOR_ST_BADREAD = 0, 1, 2
VAPOR_GT_NA = 255
VAPOR_OK, VAPOR_E_CUDA, VAPOR_E_ARG, VAPOR_E_CAPACITY, VAPOR_E_STATE = 0, -1, -2, -3, -4

# every symbol include/vapor_b200.h declares
EXPORTS = [
    "vapor_gpu_open", "vapor_gpu_close", "vapor_gpu_last_error", "vapor_gpu_set_hit_budget", "vapor_gpu_set_option",
    "vapor_gpu_score", "vapor_gpu_upload", "vapor_gpu_run", "vapor_gpu_fetch",
    "vapor_gpu_last_timings", "vapor_gpu_dotdata", "vapor_gpu_selfplot_qc", "vapor_gpu_summarize", "vapor_gpu_host_alloc", "vapor_gpu_host_free",
    "vapor_gpu_int_peak", "vapor_host_plan", "vapor_hit_mix", "vapor_b200_abi_version", "vapor_gpu_device_count",
    "vapor_gpu_pci_bus_id",
]


class VaporNativeError(RuntimeError):
    pass


class vapor_batch_t(C.Structure):
    _fields_ = [
        ("seq_bytes", C.c_void_p), ("seq_off", C.c_void_p), ("n_seq", C.c_int64),
        ("n_task", C.c_int64),
        ("task_read", C.c_void_p), ("task_ref", C.c_void_p), ("task_alt", C.c_void_p),
        ("task_miss", C.c_void_p), ("task_k", C.c_void_p), ("task_mode", C.c_void_p),
        ("n_sv", C.c_int64), ("sv_task_off", C.c_void_p),
    ]


class vapor_out_t(C.Structure):
    _fields_ = [
        ("task_score", C.c_void_p), ("task_status", C.c_void_p), ("task_stat", C.c_void_p),
        ("task_hits", C.c_void_p), ("task_hitsum", C.c_void_p),
        ("sv_qs", C.c_void_p), ("sv_gs", C.c_void_p), ("sv_gq", C.c_void_p),
        ("sv_gt", C.c_void_p), ("sv_nscore", C.c_void_p),
    ]


class vapor_timings_t(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float), ("pack_ms", C.c_float), ("tile_ms", C.c_float), ("score_ms", C.c_float),
        ("genotype_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
        ("host_prep_ms", C.c_float),
        ("cells", C.c_int64), ("hits", C.c_int64),
        ("n_plots", C.c_int64), ("n_operands", C.c_int64), ("n_strips", C.c_int64),
        ("n_waves", C.c_int64), ("n_overflow_plots", C.c_int64),
        ("launches", C.c_int64), ("bases", C.c_int64), ("padded_cells", C.c_int64),
        ("evaluated_cells", C.c_int64), ("table_ms", C.c_float), ("k2_mode", C.c_int32),
        ("table_bytes", C.c_int64), ("probe_words", C.c_int64),
        ("score_warp_ms", C.c_float), ("pad_", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def load() -> C.CDLL:
    """Load libvapor_b200.so (built by ``vapor_b200._build.build_native``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VaporNativeError(
            f"{LIB_PATH} is missing: build it with `python -m vapor_b200._build` "
            "(nvcc, sm_100a).  vapor_b200 has no CPU fallback for the scoring path.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:                                   # pragma: no cover
        raise VaporNativeError(f"cannot load {LIB_PATH}: {e}") from e
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.vapor_gpu_open.argtypes = [i32, C.POINTER(vp)]
    lib.vapor_gpu_close.argtypes = [vp]
    lib.vapor_gpu_last_error.argtypes = [vp]
    lib.vapor_gpu_last_error.restype = C.c_char_p
    lib.vapor_gpu_set_hit_budget.argtypes = [vp, i64]
    lib.vapor_gpu_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.vapor_gpu_score.argtypes = [vp, C.POINTER(vapor_batch_t), C.POINTER(vapor_out_t)]
    lib.vapor_gpu_upload.argtypes = [vp, C.POINTER(vapor_batch_t)]
    lib.vapor_gpu_run.argtypes = [vp]
    lib.vapor_gpu_fetch.argtypes = [vp, C.POINTER(vapor_out_t)]
    lib.vapor_gpu_last_timings.argtypes = [vp, C.POINTER(vapor_timings_t)]
    lib.vapor_gpu_dotdata.argtypes = [vp, i32, vp, i64, vp, i64, vp, i64, C.POINTER(i64)]
    lib.vapor_gpu_selfplot_qc.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.vapor_gpu_summarize.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, vp]
    lib.vapor_gpu_host_alloc.argtypes = [C.POINTER(vp), i64]
    lib.vapor_gpu_host_free.argtypes = [vp]
    lib.vapor_gpu_int_peak.argtypes = [vp, i32, C.POINTER(C.c_double)]
    lib.vapor_host_plan.argtypes = [C.POINTER(vapor_batch_t), i32, i32, i64, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(i64)]
    lib.vapor_hit_mix.argtypes = [C.c_uint32, C.c_uint32]
    lib.vapor_hit_mix.restype = C.c_uint64
    lib.vapor_b200_abi_version.argtypes = []
    lib.vapor_gpu_device_count.argtypes = []
    lib.vapor_gpu_pci_bus_id.argtypes = [C.c_int, C.c_char_p, C.c_int]
    lib.vapor_gpu_pci_bus_id.restype = C.c_int
    _lib = lib
    return lib
