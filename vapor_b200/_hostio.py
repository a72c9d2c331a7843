"""ctypes binding of the native region extraction (``include/vapor_hostio.h``, ``csrc/hostio.cpp``).

``FastaIndex`` answers ``samtools faidx`` queries, ``AlnFile`` + ``chop_many`` answer
``chop_pacbio_read_by_pos`` + ``minimize_pacbio_read_list`` (vapor_vali/Simple_function.pyx:339-354, 1091-1102) for
many windows per call on several host threads.  ``vapor_b200.seqio`` routes through these by default."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _native as N


class vapor_io_reads_t(C.Structure):
    _fields_ = [("n_win", C.c_int64), ("win_off", C.POINTER(C.c_int64)), ("seq_off", C.POINTER(C.c_int64)),
                ("seq_bytes", C.POINTER(C.c_uint8)), ("miss", C.POINTER(C.c_int32)), ("qname_off", C.POINTER(C.c_int64)),
                ("qname_bytes", C.POINTER(C.c_uint8)), ("n_records_seen", C.c_int64)]


EXPORTS = ["vapor_io_last_error", "vapor_io_fasta_open", "vapor_io_fasta_close", "vapor_io_fasta_fetch", "vapor_io_fasta_fetch_many",
           "vapor_io_aln_open", "vapor_io_aln_close", "vapor_io_chop_many", "vapor_io_reads_free", "vapor_io_cigar2alignstart",
           "vapor_host_scatter_runs"]
_ready = False


def lib() -> C.CDLL:
    global _ready
    L = N.load()
    if not _ready:
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        L.vapor_io_last_error.restype = C.c_char_p
        L.vapor_io_last_error.argtypes = []
        L.vapor_io_fasta_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.vapor_io_fasta_close.argtypes = [vp]
        L.vapor_io_fasta_fetch.argtypes = [vp, C.c_char_p, i64, i64, C.c_char_p, i64, C.POINTER(i64)]
        L.vapor_io_fasta_fetch_many.argtypes = [vp, i64, C.c_char_p, vp, vp, vp, i32, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.POINTER(i64))]
        L.vapor_io_aln_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.vapor_io_aln_close.argtypes = [vp]
        L.vapor_io_chop_many.argtypes = [C.POINTER(vp), i32, i64, C.c_char_p, vp, vp, vp, vp, i32, i32, C.POINTER(C.POINTER(vapor_io_reads_t))]
        L.vapor_io_reads_free.argtypes = [C.POINTER(vapor_io_reads_t)]
        L.vapor_io_cigar2alignstart.argtypes = [C.c_char_p, i64, i64, C.POINTER(i64)]
        L.vapor_host_scatter_runs.argtypes = [vp, vp, i64, vp, vp, i64, i32]
        _ready = True
    return L


def _check(rc: int, what: str):
    if rc != 0:
        raise N.VaporNativeError(f"{what} failed ({rc}): {lib().vapor_io_last_error().decode()}")


def _names(chroms: Sequence[str]) -> Tuple[bytes, np.ndarray]:
    """NUL-terminated names back to back + the offset of each (names repeat a lot: every distinct one is stored once)."""
    seen: Dict[str, int] = {}
    blob = bytearray()
    off = np.zeros(len(chroms), dtype=np.int64)
    for i, c in enumerate(chroms):
        o = seen.get(c)
        if o is None:
            o = seen[c] = len(blob)
            blob += c.encode("latin-1") + b"\0"
        off[i] = o
    return bytes(blob), off


def default_threads() -> int:
    v = os.environ.get("VAPOR_IO_THREADS")
    return int(v) if v else max(1, min(8, (os.cpu_count() or 1) // 2))


class FastaIndex:
    def __init__(self, path: str):
        self.path = path
        self._h = C.c_void_p()
        _check(lib().vapor_io_fasta_open(path.encode(), C.byref(self._h)), f"vapor_io_fasta_open({path})")

    def close(self):
        if self._h:
            lib().vapor_io_fasta_close(self._h)
            self._h = None

    def fetch(self, chrom: str, start: int, end: int) -> str:
        n = C.c_int64(0)
        cap = max(0, int(end) - int(start) + 1)
        buf = C.create_string_buffer(cap + 1)
        _check(lib().vapor_io_fasta_fetch(self._h, chrom.encode("latin-1"), int(start), int(end), buf, cap, C.byref(n)), "vapor_io_fasta_fetch")
        return buf.raw[:min(n.value, cap)].decode("latin-1")

    def fetch_many(self, regions: Sequence[Tuple[str, int, int]], threads: int = 0) -> List[str]:
        n = len(regions)
        if n == 0:
            return []
        blob, off = _names([r[0] for r in regions])
        st = np.array([int(r[1]) for r in regions], dtype=np.int64)
        en = np.array([int(r[2]) for r in regions], dtype=np.int64)
        pb, po = C.POINTER(C.c_uint8)(), C.POINTER(C.c_int64)()
        _check(lib().vapor_io_fasta_fetch_many(self._h, n, blob, off.ctypes.data, st.ctypes.data, en.ctypes.data,
                                               threads or default_threads(), C.byref(pb), C.byref(po)), "vapor_io_fasta_fetch_many")
        offs = np.ctypeslib.as_array(po, shape=(n + 1,))
        total = int(offs[-1])
        data = C.string_at(pb, total) if total else b""
        return [data[offs[i]:offs[i + 1]].decode("latin-1") for i in range(n)]


class AlnFile:
    def __init__(self, path: str):
        self.path = path
        self._h = C.c_void_p()
        _check(lib().vapor_io_aln_open(path.encode(), C.byref(self._h)), f"vapor_io_aln_open({path})")

    def close(self):
        if self._h:
            lib().vapor_io_aln_close(self._h)
            self._h = None


def chop_many(files: Sequence[AlnFile], windows: Sequence[Tuple[str, int, int, int]], max_reads: int = 20,
              threads: int = 0) -> Tuple[List[list], int]:
    """For every window ``(chrom, start, end, flank)``: the ``[[read, miss_bp, qname], ...]`` list the reference's
    ``chop_pacbio_read_by_pos`` builds over ``files`` in order, cut by ``minimize_pacbio_read_list`` when
    ``max_reads`` > 0.  Returns (lists, records seen)."""
    n = len(windows)
    if n == 0:
        return [], 0
    blob, off = _names([w[0] for w in windows])
    st = np.array([int(w[1]) for w in windows], dtype=np.int64)
    en = np.array([int(w[2]) for w in windows], dtype=np.int64)
    fl = np.array([int(w[3]) for w in windows], dtype=np.int64)
    hs = (C.c_void_p * len(files))(*[f._h for f in files])
    res = C.POINTER(vapor_io_reads_t)()
    _check(lib().vapor_io_chop_many(hs, len(files), n, blob, off.ctypes.data, st.ctypes.data, en.ctypes.data, fl.ctypes.data,
                                    int(max_reads), threads or default_threads(), C.byref(res)), "vapor_io_chop_many")
    try:
        r = res.contents
        woff = np.ctypeslib.as_array(r.win_off, shape=(n + 1,))
        nr = int(woff[-1])
        out: List[list] = [[] for _ in range(n)]
        if nr:
            soff = np.ctypeslib.as_array(r.seq_off, shape=(nr + 1,))
            qoff = np.ctypeslib.as_array(r.qname_off, shape=(nr + 1,))
            miss = np.ctypeslib.as_array(r.miss, shape=(nr,)).tolist()
            seqs = C.string_at(r.seq_bytes, int(soff[-1])).decode("latin-1")
            qn = C.string_at(r.qname_bytes, int(qoff[-1])).decode("latin-1")
            so, qo, wo = soff.tolist(), qoff.tolist(), woff.tolist()
            for w in range(n):
                out[w] = [[seqs[so[i]:so[i + 1]], miss[i], qn[qo[i]:qo[i + 1]]] for i in range(wo[w], wo[w + 1])]
        return out, int(r.n_records_seen)
    finally:
        lib().vapor_io_reads_free(res)


def cigar2alignstart(cigar: str, align_start: int, start: int) -> List[int]:
    out = (C.c_int64 * 2)()
    _check(lib().vapor_io_cigar2alignstart(cigar.encode("latin-1"), int(align_start), int(start), out), "vapor_io_cigar2alignstart")
    return [int(out[0]), int(out[1])]


def scatter_runs(dst: np.ndarray, src: np.ndarray, run_dst: np.ndarray, run_len: np.ndarray, threads: int = 2) -> None:
    """dst[run_dst[r] : run_dst[r] + run_len[r]] = the r-th run of ``src`` (runs back to back), rows of any width."""
    assert dst.flags.c_contiguous and src.flags.c_contiguous and dst.dtype == src.dtype and dst.shape[1:] == src.shape[1:]
    elem = dst.dtype.itemsize * int(np.prod(dst.shape[1:], dtype=np.int64))
    rd = np.ascontiguousarray(run_dst, dtype=np.int64); rl = np.ascontiguousarray(run_len, dtype=np.int64)
    _check(lib().vapor_host_scatter_runs(dst.ctypes.data, src.ctypes.data, elem, rd.ctypes.data, rl.ctypes.data, len(rd), threads),
           "vapor_host_scatter_runs")
