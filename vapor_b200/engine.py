"""Batching engine: packs (read, ref-structure, alt-structure) tasks of an SV set and scores
them with one C-ABI call.

The reference scores one read at a time inside each L2 driver's ``for x in all_reads:`` loop
(vapor_vali/Simple_function.pyx:1714-1726 and siblings).  Here the drivers *emit* tasks into a
``Batch``; ``Engine.score`` ships the whole batch to the GPU and returns per-task scores and
per-SV QS/GS/GT/GQ in input order.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N

MODE_ABS, MODE_W10, MODE_REDEF, MODE_ABS_AND_W10 = (N.VAPOR_MODE_ABS, N.VAPOR_MODE_W10,
                                                    N.VAPOR_MODE_REDEF, N.VAPOR_MODE_ABS_AND_W10)
GT_NAMES = ["0/0", "0/1", "1/1"]          # Simple_function.pyx:2062


def _as_u8(seq) -> np.ndarray:
    if isinstance(seq, np.ndarray):
        return np.ascontiguousarray(seq, dtype=np.uint8)
    if isinstance(seq, str):
        seq = seq.encode("latin-1")
    return np.frombuffer(bytes(seq), dtype=np.uint8)


@dataclass
class PackedBatch:
    """The arrays of ``vapor_batch_t`` (include/vapor_b200.h), numpy-owned."""
    seq_bytes: np.ndarray      # uint8
    seq_off: np.ndarray        # int64 [n_seq+1]
    task_read: np.ndarray      # int32
    task_ref: np.ndarray
    task_alt: np.ndarray
    task_miss: np.ndarray      # int32
    task_k: np.ndarray         # uint8
    task_mode: np.ndarray      # uint8
    sv_task_off: np.ndarray    # int64 [n_sv+1]

    @property
    def n_seq(self): return len(self.seq_off) - 1
    @property
    def n_task(self): return len(self.task_read)
    @property
    def n_sv(self): return len(self.sv_task_off) - 1

    def seq(self, i: int) -> bytes:
        return self.seq_bytes[self.seq_off[i]:self.seq_off[i + 1]].tobytes()

    def h2d_bytes(self) -> int:
        return int(sum(a.nbytes for a in (self.seq_bytes, self.seq_off, self.task_read, self.task_ref,
                                          self.task_alt, self.task_miss, self.task_k, self.task_mode,
                                          self.sv_task_off)))

    def validate(self):
        for name, dt in (("seq_bytes", np.uint8), ("seq_off", np.int64), ("task_read", np.int32),
                         ("task_ref", np.int32), ("task_alt", np.int32), ("task_miss", np.int32),
                         ("task_k", np.uint8), ("task_mode", np.uint8), ("sv_task_off", np.int64)):
            a = getattr(self, name)
            if a.dtype != dt or not a.flags.c_contiguous:
                setattr(self, name, np.ascontiguousarray(a, dtype=dt))
        return self

    def c_struct(self) -> N.vapor_batch_t:
        self.validate()
        b = N.vapor_batch_t()
        b.seq_bytes = self.seq_bytes.ctypes.data
        b.seq_off = self.seq_off.ctypes.data
        b.n_seq = self.n_seq
        b.n_task = self.n_task
        b.task_read = self.task_read.ctypes.data
        b.task_ref = self.task_ref.ctypes.data
        b.task_alt = self.task_alt.ctypes.data
        b.task_miss = self.task_miss.ctypes.data
        b.task_k = self.task_k.ctypes.data
        b.task_mode = self.task_mode.ctypes.data
        b.n_sv = self.n_sv
        b.sv_task_off = self.sv_task_off.ctypes.data
        return b

    def shard(self, sv_ids: Sequence[int]) -> "PackedBatch":
        """Sub-batch holding the given SVs (in the given order), sequences compacted."""
        sv_ids = np.asarray(sv_ids, dtype=np.int64)
        t0, t1 = self.sv_task_off[sv_ids], self.sv_task_off[sv_ids + 1]
        cnt = t1 - t0
        new_off = np.zeros(len(sv_ids) + 1, dtype=np.int64)
        np.cumsum(cnt, out=new_off[1:])
        tix = np.repeat(t0 - new_off[:-1], cnt) + np.arange(new_off[-1])
        used = np.unique(np.concatenate([self.task_read[tix], self.task_ref[tix], self.task_alt[tix]]))
        remap = np.full(self.n_seq, -1, dtype=np.int32)
        remap[used] = np.arange(len(used), dtype=np.int32)
        lens = (self.seq_off[1:] - self.seq_off[:-1])[used]
        soff = np.zeros(len(used) + 1, dtype=np.int64)
        np.cumsum(lens, out=soff[1:])
        src = np.repeat(self.seq_off[used] - soff[:-1], lens) + np.arange(soff[-1])
        return PackedBatch(self.seq_bytes[src], soff, remap[self.task_read[tix]], remap[self.task_ref[tix]],
                           remap[self.task_alt[tix]], self.task_miss[tix].copy(), self.task_k[tix].copy(),
                           self.task_mode[tix].copy(), new_off)


class Batch:
    """Incremental builder used by the L2 drivers."""

    def __init__(self):
        self._seqs: List[np.ndarray] = []
        self._tasks: List[tuple] = []
        self._sv_off: List[int] = [0]
        self.sv_keys: List[object] = []
        self.task_meta: List[object] = []

    def add_seq(self, seq) -> int:
        self._seqs.append(_as_u8(seq))
        return len(self._seqs) - 1

    def add_task(self, read_id: int, ref_id: int, alt_id: int, miss_bp: int, k: int, mode: int, meta=None) -> int:
        self._tasks.append((read_id, ref_id, alt_id, int(miss_bp), int(k), int(mode)))
        self.task_meta.append(meta)
        return len(self._tasks) - 1

    def end_sv(self, key=None) -> int:
        """Close the current SV: every task added since the previous ``end_sv`` belongs to it."""
        self._sv_off.append(len(self._tasks))
        self.sv_keys.append(key)
        return len(self._sv_off) - 2

    def pack(self) -> PackedBatch:
        if self._sv_off[-1] != len(self._tasks):
            self.end_sv(None)
        lens = np.array([len(s) for s in self._seqs], dtype=np.int64)
        off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        data = np.concatenate(self._seqs) if self._seqs else np.zeros(0, dtype=np.uint8)
        t = np.array(self._tasks, dtype=np.int64).reshape(-1, 6)
        return PackedBatch(np.ascontiguousarray(data, dtype=np.uint8), off,
                           t[:, 0].astype(np.int32), t[:, 1].astype(np.int32), t[:, 2].astype(np.int32),
                           t[:, 3].astype(np.int32), t[:, 4].astype(np.uint8), t[:, 5].astype(np.uint8),
                           np.array(self._sv_off, dtype=np.int64))


@dataclass
class Results:
    task_score: np.ndarray
    task_status: np.ndarray
    task_stat: np.ndarray      # [n_task, 4]
    task_hits: np.ndarray      # [n_task, 4]
    task_hitsum: np.ndarray    # [n_task, 4] uint64
    sv_qs: np.ndarray
    sv_gs: np.ndarray
    sv_gq: np.ndarray
    sv_gt: np.ndarray          # uint8, 255 = NA
    sv_nscore: np.ndarray

    def d2h_bytes(self) -> int:
        return int(sum(getattr(self, f).nbytes for f in self.__dataclass_fields__))

    def sv_scores(self, batch: PackedBatch, s: int) -> List[float]:
        """vapor_score_list of SV ``s`` as the reference driver would have returned it."""
        t0, t1 = int(batch.sv_task_off[s]), int(batch.sv_task_off[s + 1])
        ok = self.task_status[t0:t1] == N.VAPOR_ST_SCORED
        return [float(v) for v in self.task_score[t0:t1][ok]]


def _alloc_results(n_task: int, n_sv: int) -> Results:
    return Results(np.zeros(n_task, np.float64), np.zeros(n_task, np.uint8), np.zeros((n_task, 4), np.float64),
                   np.zeros((n_task, 4), np.uint32), np.zeros((n_task, 4), np.uint64),
                   np.zeros(n_sv, np.float64), np.zeros(n_sv, np.float64), np.zeros(n_sv, np.float64),
                   np.full(n_sv, 255, np.uint8), np.zeros(n_sv, np.int32))


def _out_struct(r: Results) -> N.vapor_out_t:
    o = N.vapor_out_t()
    for f in r.__dataclass_fields__:
        setattr(o, f, getattr(r, f).ctypes.data)
    return o


class Engine:
    """One handle on one CUDA device (``vapor_gpu_open``)."""

    def __init__(self, device: int = 0, hit_budget_bytes: int = 0):
        self._lib = N.load()
        self._h = C.c_void_p()
        rc = self._lib.vapor_gpu_open(int(device), C.byref(self._h))
        if rc != 0:
            msg = self._lib.vapor_gpu_last_error(None).decode()
            self._h = None
            raise N.VaporNativeError(f"vapor_gpu_open(device={device}) failed ({rc}): {msg}")
        self.device = device
        self._n_task = self._n_sv = 0
        if hit_budget_bytes:
            self._lib.vapor_gpu_set_hit_budget(self._h, int(hit_budget_bytes))

    # -- lifecycle -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            for p in getattr(self, "_pinned", []):
                self._lib.vapor_gpu_host_free(p)
            self._pinned = []
            self._lib.vapor_gpu_close(self._h)
            self._h = None

    def __enter__(self): return self
    def __exit__(self, *a): self.close()
    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.vapor_gpu_last_error(self._h).decode()
            if "BADREAD" in msg:
                raise KeyError(msg)        # the reference raises KeyError from invert_base (Simple_function.pyx:1421)
            raise N.VaporNativeError(f"{what} failed ({rc}): {msg}")

    def set_option(self, name: str, value: int):
        """Named tunables of the library (``vapor_gpu_set_option``); none changes a result."""
        self._check(self._lib.vapor_gpu_set_option(self._h, name.encode(), int(value)), f"vapor_gpu_set_option({name})")

    # -- scoring ---------------------------------------------------------------------------
    def score(self, batch: PackedBatch) -> Results:
        """Blocking: host buffers in, host buffers out (``vapor_gpu_score``)."""
        b = batch.c_struct()
        res = _alloc_results(batch.n_task, batch.n_sv)
        o = _out_struct(res)
        self._check(self._lib.vapor_gpu_score(self._h, C.byref(b), C.byref(o)), "vapor_gpu_score")
        return res

    def upload(self, batch: PackedBatch):
        b = batch.c_struct()
        self._check(self._lib.vapor_gpu_upload(self._h, C.byref(b)), "vapor_gpu_upload")
        self._n_task, self._n_sv = batch.n_task, batch.n_sv

    def run(self):
        self._check(self._lib.vapor_gpu_run(self._h), "vapor_gpu_run")

    def fetch(self, into: Optional[Results] = None) -> Results:
        res = into if into is not None else _alloc_results(self._n_task, self._n_sv)
        o = _out_struct(res)
        self._check(self._lib.vapor_gpu_fetch(self._h, C.byref(o)), "vapor_gpu_fetch")
        return res

    def timings(self) -> dict:
        t = N.vapor_timings_t()
        self._check(self._lib.vapor_gpu_last_timings(self._h, C.byref(t)), "vapor_gpu_last_timings")
        return t.as_dict()

    # -- single plot -------------------------------------------------------------------------
    def dotdata(self, k: int, seq1, seq2) -> np.ndarray:
        """``dotdata(kmerlen, seq1=read, seq2=structure)`` (Simple_function.pyx:545-549): ``(H, 2)`` int32
        rows ``(x, y)`` in the reference's list order."""
        r, s = _as_u8(seq1), _as_u8(seq2)
        n = C.c_int64(0)
        cap = max(1024, 2 * (len(r) + len(s)))
        while True:
            xy = np.zeros((cap, 2), dtype=np.int32)
            rc = self._lib.vapor_gpu_dotdata(self._h, int(k), r.ctypes.data, len(r), s.ctypes.data, len(s),
                                             xy.ctypes.data, cap, C.byref(n))
            self._check(rc, "vapor_gpu_dotdata")
            if n.value <= cap:
                return xy[:n.value]
            cap = int(n.value)

    def selfplot_qc(self, seqs, ks) -> np.ndarray:
        """Self-plot counts for ``window_size_refine`` (``vapor_gpu_selfplot_qc``): one row per sequence,
        columns H, diag, lower, min x, max x, min y, max y, status."""
        arrs = [_as_u8(s) for s in seqs]
        n = len(arrs)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(a) for a in arrs], out=off[1:])
        data = np.ascontiguousarray(np.concatenate(arrs) if n and off[-1] else np.zeros(0, np.uint8), dtype=np.uint8)
        kk = np.ascontiguousarray(np.asarray(ks, dtype=np.uint8).reshape(-1))
        if len(kk) != n:
            raise ValueError("one k per sequence")
        out = np.zeros((n, 8), dtype=np.int64)
        rc = self._lib.vapor_gpu_selfplot_qc(self._h, data.ctypes.data if len(data) else None, off.ctypes.data, n,
                                             kk.ctypes.data if n else None, out.ctypes.data if n else None)
        self._check(rc, "vapor_gpu_selfplot_qc")
        return out

    def summarize(self, score_lists):
        """QS/GS/GT/GQ for a list of per-SV score lists (kernel 4 alone, ``vapor_gpu_summarize``).
        Returns arrays (qs, gs, gq, gt, nscore); gt == 255 marks the reference's 'NA' row."""
        n_sv = len(score_lists)
        off = np.zeros(n_sv + 1, dtype=np.int64)
        np.cumsum([len(s) for s in score_lists], out=off[1:])
        flat = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.float64) for s in score_lists])
                                    if n_sv and off[-1] else np.zeros(0), dtype=np.float64)
        qs = np.zeros(n_sv); gs = np.zeros(n_sv); gq = np.zeros(n_sv)
        gt = np.full(n_sv, 255, np.uint8); ns = np.zeros(n_sv, np.int32)
        rc = self._lib.vapor_gpu_summarize(self._h, flat.ctypes.data if len(flat) else None, off.ctypes.data, n_sv,
                                           qs.ctypes.data, gs.ctypes.data, gq.ctypes.data, gt.ctypes.data, ns.ctypes.data)
        self._check(rc, "vapor_gpu_summarize")
        return qs, gs, gq, gt, ns

    # -- pinned host staging ---------------------------------------------------------------
    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array backed by page-locked host memory (``vapor_gpu_host_alloc``); freed on close()."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        nbytes = max(1, n * dtype.itemsize)
        p = C.c_void_p()
        rc = self._lib.vapor_gpu_host_alloc(C.byref(p), nbytes)
        if rc != 0:
            raise N.VaporNativeError(f"vapor_gpu_host_alloc({nbytes}) failed ({rc})")
        if not hasattr(self, "_pinned"):
            self._pinned = []
        self._pinned.append(p)
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def pin_batch(self, batch: PackedBatch) -> PackedBatch:
        """Copy a batch into pinned host memory so H2D copies run at full PCIe speed."""
        out = {}
        for f in batch.__dataclass_fields__:
            a = getattr(batch, f)
            b = self.pinned_empty(a.shape, a.dtype)
            b[...] = a
            out[f] = b
        return PackedBatch(**out)

    def pinned_results(self, n_task: int, n_sv: int) -> Results:
        r = _alloc_results(n_task, n_sv)
        for f in r.__dataclass_fields__:
            a = getattr(r, f)
            b = self.pinned_empty(a.shape, a.dtype)
            b[...] = a
            setattr(r, f, b)
        return r

    def score_into(self, batch: PackedBatch, res: Results) -> Results:
        """``score`` writing into caller-provided (e.g. pinned) result arrays."""
        b = batch.c_struct()
        o = _out_struct(res)
        self._check(self._lib.vapor_gpu_score(self._h, C.byref(b), C.byref(o)), "vapor_gpu_score")
        return res

    def int_peak(self, which: int = 0) -> float:
        v = C.c_double(0)
        self._check(self._lib.vapor_gpu_int_peak(self._h, int(which), C.byref(v)), "vapor_gpu_int_peak")
        return v.value


class Pipeline:
    """Double-buffered scoring of a stream of batches on one GPU: ``depth`` handles (``vapor_gpu_open`` each, own
    stream and device buffers), one worker thread per handle.  Every batch goes through the three public steps
    ``vapor_gpu_upload`` (host planning + H2D), ``vapor_gpu_run`` (kernels), ``vapor_gpu_fetch`` (D2H); one lock per
    step keeps the handles out of each other's way, so while one handle's kernels run the next batch is planned and
    copied in on the other: the steps form a pipeline whose beat is the slowest step instead of their sum
    (the C-ABI calls release the GIL).  Results come back in submission order.

        with Pipeline(device=0, depth=2) as pipe:
            for res in pipe.map(batches): ...
    """

    def __init__(self, device: int = 0, depth: int = 2, options: Optional[dict] = None):
        self.engines = [Engine(device) for _ in range(max(1, depth))]
        for e in self.engines:
            for k, v in (options or {}).items():
                e.set_option(k, v)

    def close(self):
        for e in self.engines:
            e.close()

    def __enter__(self): return self
    def __exit__(self, *a): self.close()

    def map(self, batches, results: Optional[Sequence[Optional[Results]]] = None, after=None):
        """Score every batch; ``results[i]`` (optional) is a caller-provided (e.g. pinned) Results to fill.
        ``after(results_i)`` (optional) runs in the worker thread right after batch i was scored -- e.g. the scatter
        of a shard's results into shared input-order arrays -- while the other handle keeps the GPU busy."""
        import threading
        import time
        batches = list(batches)
        n = len(batches)
        self.stage_s = {}
        out: List[Optional[Results]] = [None] * n
        errs: List[BaseException] = []
        nxt = [0]
        lock = threading.Lock()
        copy_in, kernels, copy_out = threading.Lock(), threading.Lock(), threading.Lock()

        def work(eng):
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= n or errs:
                    return
                try:
                    into = results[i] if results is not None and results[i] is not None else _alloc_results(batches[i].n_task, batches[i].n_sv)
                    t0 = time.perf_counter()
                    with copy_in:
                        t1 = time.perf_counter()
                        eng.upload(batches[i])
                        t2 = time.perf_counter()
                    with kernels:
                        t3 = time.perf_counter()
                        eng.run()
                        t4 = time.perf_counter()
                    with copy_out:
                        t5 = time.perf_counter()
                        out[i] = eng.fetch(into)
                        t6 = time.perf_counter()
                    if after is not None:
                        after(out[i])
                    t7 = time.perf_counter()
                    with lock:                           # wall seconds per stage, summed over the batches (self.stage_s)
                        for k_, v_ in (("wait_upload", t1 - t0), ("upload", t2 - t1), ("wait_run", t3 - t2), ("run", t4 - t3),
                                       ("wait_fetch", t5 - t4), ("fetch", t6 - t5), ("after", t7 - t6)):
                            self.stage_s[k_] = self.stage_s.get(k_, 0.0) + v_
                except BaseException as e:          # noqa: BLE001
                    errs.append(e)
                    return
        th = [threading.Thread(target=work, args=(e,)) for e in self.engines]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return out


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device: int, sysfs: str = "/sys") -> Optional[int]:
    """Restrict this process to the CPUs of the NUMA node GPU ``device`` hangs off, so that the pinned buffers it
    allocates afterwards are first-touched on that node and its H2D copies do not cross the socket interconnect (with one
    process per GPU, eight ranks otherwise pull their inputs through whichever node they happened to start on).
    Returns the node, or None when the topology is not visible (single node, container without sysfs, no GPU): a no-op
    then.  Reads ``<sysfs>/bus/pci/devices/<bus id>/numa_node`` and ``<sysfs>/devices/system/node/node<N>/cpulist``."""
    import os
    try:
        buf = C.create_string_buffer(32)
        if N.load().vapor_gpu_pci_bus_id(int(device), buf, 32) != 0:
            return None
        bus = buf.value.decode().lower()
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bus, "numa_node")).read().strip())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(os.path.join(sysfs, "devices/system/node", f"node{node}", "cpulist")).read())
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError):
        return None


def host_plan(batch: PackedBatch, k2_mode: int = 1, threads: int = 0, wave_budget_bytes: int = 0) -> dict:
    """``vapor_host_plan``: plan a batch on the host only (no GPU needed): wall ms, plan digest, counts."""
    lib = N.load()
    b = batch.c_struct()
    ms, dg = C.c_double(0), C.c_uint64(0)
    cnt = (C.c_int64 * 8)()
    rc = lib.vapor_host_plan(C.byref(b), int(k2_mode), int(threads), int(wave_budget_bytes), C.byref(ms), C.byref(dg), cnt)
    if rc != 0:
        raise N.VaporNativeError(f"vapor_host_plan failed ({rc}): {lib.vapor_gpu_last_error(None).decode()}")
    names = ("operands", "plots", "tasks", "waves", "table_chunks", "join_items", "cells", "max_wave_hits")
    return {"ms": ms.value, "digest": dg.value, **{n: int(v) for n, v in zip(names, cnt)}}


def hit_mix(x, y) -> np.ndarray:
    """numpy twin of ``vapor_hit_mix`` (include/vapor_b200.h) for checksum comparisons."""
    x = np.asarray(x, dtype=np.uint64)
    y = np.asarray(y, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (x << np.uint64(32)) | y
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def hit_checksum(dots: np.ndarray) -> int:
    if len(dots) == 0:
        return 0
    with np.errstate(over="ignore"):
        return int(np.sum(hit_mix(dots[:, 0], dots[:, 1]), dtype=np.uint64))
