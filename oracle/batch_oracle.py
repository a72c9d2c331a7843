"""Score a whole ``PackedBatch`` with the CPU oracle, task by task.  TEST INFRASTRUCTURE ONLY
(used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs)."""
from __future__ import annotations

import numpy as np

from . import vapor_oracle as O


def _hit_mix(x, y):
    x = np.asarray(x, dtype=np.uint64); y = np.asarray(y, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (x << np.uint64(32)) | y
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def hit_checksum(dots) -> int:
    if len(dots) == 0:
        return 0
    with np.errstate(over="ignore"):
        return int(np.sum(_hit_mix(dots[:, 0], dots[:, 1]), dtype=np.uint64))


def score_task(read: str, ref: str, alt: str, miss: int, k: int, mode: int, impl=O):
    """Returns (score or None, stat[4], hits[4], hitsum[4]) for one task.  ``impl`` is the oracle
    module or the reference module itself (same function names)."""
    x = [read, miss, "q"]
    stat = [0.0] * 4
    fa = {O.MODE_ABS: "calcu_vapor_single_read_score_abs_dis_m1b",
          O.MODE_W10: "calcu_vapor_single_read_score_within_10Perc_m1b",
          O.MODE_REDEF: "calcu_vapor_single_read_score_directed_dis_m1b_redefine_diagnal",
          O.MODE_ABS_AND_W10: "calcu_vapor_single_read_score_abs_dis_m1b"}[mode]
    pa = getattr(impl, fa)(ref, alt, x, k)
    stat[0], stat[1] = float(pa[0]), float(pa[1])
    sa = None if 0 in pa else 1 - float(pa[1]) / float(pa[0])
    score = sa
    if mode == O.MODE_ABS_AND_W10:
        pb = impl.calcu_vapor_single_read_score_within_10Perc_m1b(ref, alt, x, k)
        stat[2], stat[3] = float(pb[0]), float(pb[1])
        sb = None if 0 in pb else 1 - float(pb[1]) / float(pb[0])
        if sa is not None and sb is not None:
            score = min([sa, sb])
        elif sa is None:
            score = sb
    return score, stat


def task_hit_info(read: str, ref: str, alt: str, miss: int, k: int, mode: int):
    """len(dotdata) and coordinate checksum of the plots a task evaluates (oracle dotdata)."""
    up = mode in (O.MODE_ABS, O.MODE_ABS_AND_W10)
    ra, aa = (ref.upper(), alt.upper()) if up else (ref, alt)
    plots = [O.dotdata(k, read, ra[miss:]), O.dotdata(k, read, aa[miss:])]
    if mode == O.MODE_ABS_AND_W10:
        plots += [O.dotdata(k, read, ref[miss:]), O.dotdata(k, read, alt[miss:])]
    hits = [len(p) for p in plots] + [0] * (4 - len(plots))
    sums = [hit_checksum(p) for p in plots] + [0] * (4 - len(plots))
    return hits, sums


def score_batch(batch, impl=O, with_hits: bool = True, task_range=None):
    """Oracle results for every task/SV of a PackedBatch (same layout as vapor_b200.engine.Results)."""
    nt, nsv = batch.n_task, batch.n_sv
    score = np.zeros(nt); status = np.zeros(nt, np.uint8)
    stat = np.zeros((nt, 4)); hits = np.zeros((nt, 4), np.uint32); hitsum = np.zeros((nt, 4), np.uint64)
    seqs = {}

    def seq(i):
        if i not in seqs:
            seqs[i] = batch.seq(int(i)).decode("latin-1")
        return seqs[i]
    rng = range(nt) if task_range is None else task_range
    for t in rng:
        read, ref, alt = seq(batch.task_read[t]), seq(batch.task_ref[t]), seq(batch.task_alt[t])
        miss, k, mode = int(batch.task_miss[t]), int(batch.task_k[t]), int(batch.task_mode[t])
        try:
            sc, st = score_task(read, ref, alt, miss, k, mode, impl)
        except KeyError:
            status[t] = 2
            continue
        stat[t] = st
        if sc is not None:
            score[t] = sc; status[t] = 1
        if with_hits:
            h, s = task_hit_info(read, ref, alt, miss, k, mode)
            hits[t] = h; hitsum[t] = np.array(s, dtype=np.uint64)
    qs = np.zeros(nsv); gs = np.zeros(nsv); gq = np.zeros(nsv); gt = np.full(nsv, 255, np.uint8)
    nscore = np.zeros(nsv, np.int32); rec = [None] * nsv
    for s in range(nsv):
        t0, t1 = int(batch.sv_task_off[s]), int(batch.sv_task_off[s + 1])
        sc = [float(score[t]) for t in range(t0, t1) if status[t] == 1]
        nscore[s] = len(sc)
        summ = O.summarize_sv(sc)
        if summ is not None:
            qs[s], gs[s], gq[s], gt[s], rec[s] = summ["QS"], summ["GS"], summ["GQ"], summ["GT"], summ["Rec"]
    return dict(task_score=score, task_status=status, task_stat=stat, task_hits=hits, task_hitsum=hitsum,
                sv_qs=qs, sv_gs=gs, sv_gq=gq, sv_gt=gt, sv_nscore=nscore, sv_rec=rec)
