"""SV sharding across GPUs: independent per-SV work queues, no data-path collective.

Every SV (indeed every read) is independent, so N GPUs means N sub-batches.  The SV
list is partitioned by a greedy longest-processing-time rule on the recurrence cells of each SV, every
part keeps its SVs in input order, and results are written back into arrays indexed by the original
task / SV position, so the output order is the input order whatever N is.
"""
from __future__ import annotations

import threading
from typing import Callable, List, Optional, Sequence

import numpy as np

from .engine import PackedBatch, Results, _alloc_results


def sv_costs(batch: PackedBatch) -> np.ndarray:
    """Recurrence cells per SV: sum over its reads of n * (m_ref + m_alt)."""
    lens = batch.seq_off[1:] - batch.seq_off[:-1]
    k = batch.task_k.astype(np.int64)
    miss = batch.task_miss.astype(np.int64)
    n = np.maximum(lens[batch.task_read] - k + 1, 0)
    m = (np.maximum(lens[batch.task_ref] - miss - k + 1, 0) + np.maximum(lens[batch.task_alt] - miss - k + 1, 0))
    cells = n * m
    csum = np.concatenate([[0], np.cumsum(cells)])
    return csum[batch.sv_task_off[1:]] - csum[batch.sv_task_off[:-1]]


def partition_svs(costs: Sequence[int], n_parts: int) -> List[np.ndarray]:
    """Greedy LPT: heaviest SV first onto the lightest part; each part returned in ascending SV order."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(n_parts, dtype=np.int64)
    parts: List[List[int]] = [[] for _ in range(n_parts)]
    for s in order:
        p = int(np.argmin(load))
        parts[p].append(int(s))
        load[p] += int(costs[s]) + 1
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def part_task_index(sv_task_off: np.ndarray, sv_ids: np.ndarray) -> np.ndarray:
    """Input positions of the tasks of the SVs ``sv_ids`` (in that order) in the whole list."""
    sv_ids = np.asarray(sv_ids, dtype=np.int64)
    t0, t1 = sv_task_off[sv_ids], sv_task_off[sv_ids + 1]
    cnt = t1 - t0
    loc = np.zeros(len(sv_ids) + 1, dtype=np.int64)
    np.cumsum(cnt, out=loc[1:])
    return np.repeat(t0 - loc[:-1], cnt) + np.arange(loc[-1])


def scatter_part(out: Results, tix: np.ndarray, sv_ids: np.ndarray, part: Results, sv_task_off: Optional[np.ndarray] = None) -> None:
    """Write one part's results to their input positions in ``out`` (pre-sized arrays of the whole list).  With
    ``sv_task_off`` (task offsets of the whole list) the copy is one native memcpy per SV (``vapor_host_scatter_runs``);
    without it, numpy fancy indexing with the task positions ``tix``."""
    if len(sv_ids) == 0:
        return
    sv_ids = np.asarray(sv_ids, dtype=np.int64)
    if sv_task_off is not None:
        from . import _hostio
        run_dst = np.ascontiguousarray(sv_task_off[sv_ids], dtype=np.int64)
        run_len = np.ascontiguousarray(sv_task_off[sv_ids + 1] - sv_task_off[sv_ids], dtype=np.int64)
        for f in ("task_score", "task_status", "task_stat", "task_hits", "task_hitsum"):
            _hostio.scatter_runs(getattr(out, f), np.ascontiguousarray(getattr(part, f)), run_dst, run_len)
    else:
        for f in ("task_score", "task_status", "task_stat", "task_hits", "task_hitsum"):
            getattr(out, f)[tix] = getattr(part, f)
    for f in ("sv_qs", "sv_gs", "sv_gq", "sv_gt", "sv_nscore"):
        getattr(out, f)[sv_ids] = getattr(part, f)


def merge_results(batch: PackedBatch, parts: Sequence[np.ndarray], results: Sequence[Results]) -> Results:
    """Scatter per-part results back to input order."""
    out = _alloc_results(batch.n_task, batch.n_sv)
    for sv_ids, r in zip(parts, results):
        if len(sv_ids):
            scatter_part(out, None, np.asarray(sv_ids, dtype=np.int64), r, sv_task_off=batch.sv_task_off)
    return out


class SharedResults:
    """The result arrays of a whole SV list in shared host memory (files under /dev/shm mapped by every process):
    the pre-sized host arrays indexed by input position of SURVEY.md 8(e).  One process per GPU scores its part
    and writes it straight to its input positions (``scatter_part``); after a barrier the owner (rank 0) holds
    the gathered results in input order.  No collective, no pickling, no copy through a socket."""

    SPEC = (("task_score", np.float64, 1), ("task_status", np.uint8, 1), ("task_stat", np.float64, 4),
            ("task_hits", np.uint32, 4), ("task_hitsum", np.uint64, 4), ("sv_qs", np.float64, 0),
            ("sv_gs", np.float64, 0), ("sv_gq", np.float64, 0), ("sv_gt", np.uint8, 0), ("sv_nscore", np.int32, 0))

    def __init__(self, tag: str, n_task: int, n_sv: int, create: bool, directory: Optional[str] = None):
        import os
        import tempfile
        d = directory or ("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
        self.paths, arrs = [], {}
        for name, dt, width in self.SPEC:
            n = n_task if name.startswith("task_") else n_sv
            shape = (n, width) if width > 1 else (n,)
            path = os.path.join(d, f"vapor_b200_{tag}_{name}.bin")
            self.paths.append(path)
            arrs[name] = np.memmap(path, dtype=dt, mode="w+" if create else "r+", shape=shape) if n else np.zeros(shape, dt)
        if create and n_sv:
            arrs["sv_gt"][:] = 255
        self.results = Results(**arrs)
        self.owner = create

    def close(self):
        import os
        self.results = None
        if self.owner:
            for p in self.paths:
                try:
                    os.unlink(p)
                except OSError:
                    pass


def score_sharded(batch: PackedBatch, scorers: Sequence[Callable[[PackedBatch], Results]]) -> Results:
    """One process, one scorer (= one Engine.score bound to one GPU) per part, driven from threads
    (the C-ABI call releases the GIL).  Output is in input order."""
    n = len(scorers)
    parts = partition_svs(sv_costs(batch), n)
    res: List[Optional[Results]] = [None] * n
    errs: List[BaseException] = []

    def work(i):
        try:
            res[i] = scorers[i](batch.shard(parts[i])) if len(parts[i]) else _alloc_results(0, 0)
        except BaseException as e:      # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]
    return merge_results(batch, parts, res)


def score_distributed(batch: PackedBatch, scorer: Callable[[PackedBatch], Results], rank: int, world: int,
                      dist=None) -> Optional[Results]:
    """One process per GPU (torch.distributed plumbing only): every rank scores its own part, rank 0
    gathers the per-part results (python objects, off the data path) and returns them in input order."""
    parts = partition_svs(sv_costs(batch), world)
    mine = scorer(batch.shard(parts[rank])) if len(parts[rank]) else _alloc_results(0, 0)
    if world == 1 or dist is None:
        return merge_results(batch, parts, [mine])
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None
    return merge_results(batch, parts, gathered)
