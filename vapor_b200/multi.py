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


def merge_results(batch: PackedBatch, parts: Sequence[np.ndarray], results: Sequence[Results]) -> Results:
    """Scatter per-part results back to input order."""
    out = _alloc_results(batch.n_task, batch.n_sv)
    for sv_ids, r in zip(parts, results):
        if len(sv_ids) == 0:
            continue
        t0, t1 = batch.sv_task_off[sv_ids], batch.sv_task_off[sv_ids + 1]
        cnt = t1 - t0
        loc = np.concatenate([[0], np.cumsum(cnt)])
        tix = np.repeat(t0 - loc[:-1], cnt) + np.arange(loc[-1])
        for f in ("task_score", "task_status", "task_stat", "task_hits", "task_hitsum"):
            getattr(out, f)[tix] = getattr(r, f)
        for f in ("sv_qs", "sv_gs", "sv_gq", "sv_gt", "sv_nscore"):
            getattr(out, f)[sv_ids] = getattr(r, f)
    return out


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
