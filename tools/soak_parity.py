#!/usr/bin/env python
"""One-off soak: the CUDA path against the oracle on a few thousand tasks of every SV type, k in {10,20,30,40},
miss_bp > 0 and soft-masked stretches.  Prints one JSON line with the counts of compared and differing items.

    python tools/soak_parity.py [--n-sv 240] [--procs 16]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def _oracle_chunk(args):
    from oracle import batch_oracle as BO
    batch, lo, hi = args
    e = BO.score_batch(batch, task_range=range(lo, hi))
    return lo, hi, {k: e[k][lo:hi] for k in ("task_score", "task_status", "task_stat", "task_hits", "task_hitsum")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-sv", type=int, default=240)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--recipe", default="simple", choices=["simple", "complex", "large"])
    ap.add_argument("--k2-mode", type=int, default=1, help="1 = join kernel, 0 = all-pairs tile kernel, 2 = both (must agree)")
    a = ap.parse_args()
    from vapor_b200 import synth
    from vapor_b200.engine import Engine
    from oracle import vapor_oracle as O
    if a.recipe == "simple":
        w = synth.make_workload(a.n_sv, seed=4242, size_range=(50, 5000), reads_per_sv=20, max_miss=4, lowercase_every=7,
                                k_choices=(10, 10, 10, 20, 10, 30, 10, 40))
    elif a.recipe == "complex":        # BASELINE config 3: every complex kind incl. 2-3-allele records and junction fallbacks
        w = synth.make_workload(a.n_sv, seed=4343, recipe="complex", size_range=(200, 5000), reads_per_sv=20,
                                k_choices=(10, 10, 10, 20, 10, 10, 30))
    else:                              # BASELINE config 4: 10-60 kb windows, every k
        w = synth.make_workload(a.n_sv, seed=4444, recipe="large", size_range=(10000, 60000), reads_per_sv=4, k_choices=(10, 20, 30, 40))
    t0 = time.time()
    with Engine(0) as eng:
        eng.set_option("k2_mode", 0 if a.k2_mode == 0 else 1)
        res = eng.score(w.batch)
        if a.k2_mode == 2:
            eng.set_option("k2_mode", 0)
            other = eng.score(w.batch)
            for f in res.__dataclass_fields__:
                assert np.array_equal(getattr(res, f), getattr(other, f)), f"kernel-2 variants differ in {f}"
        eng.set_option("k2_mode", 0 if a.k2_mode == 0 else 1)
        eng.set_option("k3_mode", 0)                      # the CTA-per-task kernel 3 must agree with the warp-per-task one
        other = eng.score(w.batch)
        for f in res.__dataclass_fields__:
            assert np.array_equal(getattr(res, f), getattr(other, f)), f"kernel-3 variants differ in {f}"
    t_gpu = time.time() - t0
    nt = w.batch.n_task
    step = max(1, nt // (a.procs * 8))
    jobs = [(w.batch, lo, min(nt, lo + step)) for lo in range(0, nt, step)]
    t0 = time.time()
    with mp.get_context("fork").Pool(a.procs) as pool:
        parts = pool.map(_oracle_chunk, jobs)
    t_cpu = time.time() - t0
    diff = {k: 0 for k in ("task_score", "task_status", "task_stat", "task_hits", "task_hitsum")}
    for lo, hi, e in parts:
        for k in diff:
            diff[k] += int(np.sum(np.any(np.asarray(getattr(res, k)[lo:hi]).reshape(hi - lo, -1) != np.asarray(e[k]).reshape(hi - lo, -1), axis=1)))
    # per-SV summaries from the GPU scores through the oracle's summary
    gt_diff = qs_diff = gq_diff = 0
    for s in range(w.batch.n_sv):
        sc = res.sv_scores(w.batch, s)
        summ = O.summarize_sv(sc)
        if summ is None:
            gt_diff += int(res.sv_gt[s] != 255)
            continue
        gt_diff += int(res.sv_gt[s] != summ["GT"])
        qs_diff += int(abs(res.sv_qs[s] - summ["QS"]) > 1e-5 or abs(res.sv_gs[s] - summ["GS"]) > 1e-5)
        gq_diff += int(abs(res.sv_gq[s] - summ["GQ"]) > 1e-3)
    print(json.dumps({"recipe": a.recipe, "k2_mode": {0: "tile", 1: "join", 2: "join and tile (identical)"}[a.k2_mode],
                      "sv_types": sorted(set(w.sv_type)), "modes": sorted(set(int(m) for m in w.batch.task_mode)),
                      "k": sorted(set(int(k) for k in w.batch.task_k)), "tasks": nt, "svs": w.batch.n_sv, "scored": int((res.task_status == 1).sum()),
                      "tasks_differing": diff, "sv_gt_differing": gt_diff, "sv_qs_gs_beyond_1e-5": qs_diff, "sv_gq_beyond_1e-3": gq_diff,
                      "gpu_s": round(t_gpu, 2), "oracle_s_on_%d_procs" % a.procs: round(t_cpu, 1)}))


if __name__ == "__main__":
    main()
