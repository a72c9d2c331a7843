#!/usr/bin/env python
"""bench.py -- reads scored/s of the VaPoR per-read scoring path on N B200s (one JSON line).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3|4|5] [--n-sv M] [--weak]

Workloads (BASELINE.json `configs`, seeded: SV i derives from default_rng([20261018 + config index, i])):
  --config 5 (default)  configs[4]: one genome-scale 100 000-SV list (DEL/TANDUP/INV/INS cycled, 50 bp-5 kb, 20
                        CLR-like reads per SV at 15 % error = 30x coverage after the reference's 20-read cap, k = 10),
                        SV-sharded over the N GPUs -- the configuration the north_star's target is quoted on.
  --config 2            configs[1]: 10 000 simple SVs of the same kind.
  --config 3            configs[2]: 10 000 complex events (DEL_INV, DUP_INV, DISDUP, DEL_DUP_INV, multi-allele
                        `Other=` records with 2-3 alternative haplotypes, junction fallbacks; modes ABS/REDEF/W10 mixed).
  --config 4            configs[3]: 400 large events, 10-100 kb windows, k cycled over 10/20/30/40.
A *step* is one pass of the hot path (kernels 1-4) over the whole SV list.

Multi-GPU (torchrun, one process per GPU): "scaling": "strong".  Every rank computes the cost of every SV of the list
(vapor_b200.synth.workload_costs, no sequence needed), takes its part from the product's partitioner
(vapor_b200.multi.partition_svs: greedy longest-processing-time on recurrence cells), builds and scores only that
part, and writes its results to their input positions in result arrays shared by all ranks
(vapor_b200.multi.SharedResults, /dev/shm): after the barrier rank 0 holds the whole list's results in input order.
No data-path collective; NCCL carries the barrier and the max-over-ranks of the timings only.  `output_checksum`
(SHA-256 over the gathered hit checksums, scores, statuses and genotype calls) must be identical at N = 1/2/4/8.
--weak gives every rank its own M SVs instead (SVs rank*M ...), as round 1 measured.

  value  : reads scored/s with every rank's part resident in HBM (vapor_gpu_upload once, vapor_gpu_run per step),
           timed with CUDA events on the library's stream, max over ranks.
  e2e    : the same metric through the public call (vapor_gpu_score on pinned HOST buffers: host planning + H2D +
           kernels + D2H every step) plus the gather into input order, wall clock between barriers, max over ranks.
  roofline: the dominant kernel of the step (by CUDA-event time) against its bound; `kernels` lists all of them.
  cpu_baseline / --impl reference: the reference's own code (oracle/_ref = Cython build of the untouched
           Simple_function.pyx when it travelled with the snapshot, else the numpy oracle port) on all host cores, on a
           stratified sample of the same SV list.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED0 = 20261018               # seed = 20261018 + config index
READS_PER_SV = 20
METRIC = "reads_scored_per_sec"
UNIT = "reads/s"

CONFIGS = {
    2: dict(index=1, recipe="simple", n_sv=10000, size_range=(50, 5000), k=(10,),
            name="BASELINE configs[1]: 10k simple SVs DEL/TANDUP/INV/INS 50bp-5kb, 20 CLR-like reads (15% error) per SV, k=10"),
    3: dict(index=2, recipe="complex", n_sv=10000, size_range=(200, 5000), k=(10,),
            name="BASELINE configs[2]: 10k complex events (DEL_INV, DUP_INV, DISDUP, DEL_DUP_INV, 2-3-allele Other= records, "
                 "junction fallbacks), 20 CLR-like reads per event scored against every alternative haplotype, modes ABS/REDEF/W10, k=10"),
    4: dict(index=3, recipe="large", n_sv=400, size_range=(10000, 100000), k=(10, 20, 30, 40),
            name="BASELINE configs[3]: 400 large events (DEL/TANDUP/INV/INS) with 10-100 kb windows, 20 CLR-like reads each, "
                 "k cycled over 10/20/30/40"),
    5: dict(index=4, recipe="simple", n_sv=100000, size_range=(50, 5000), k=(10,),
            name="BASELINE configs[4]: genome-scale 100k-SV list (DEL/TANDUP/INV/INS 50bp-5kb), 20 CLR-like reads (15% error) "
                 "per SV = 30x after the 20-read cap, k=10, SV-sharded across the GPUs"),
}


def _dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _cfg(args):
    c = dict(CONFIGS[args.config])
    if args.n_sv > 0:
        c["n_sv"] = args.n_sv
    c["seed"] = SEED0 + c["index"]
    return c


def _gen_kwargs(c):
    return dict(seed=c["seed"], recipe=c["recipe"], size_range=c["size_range"], reads_per_sv=READS_PER_SV, k_choices=c["k"])


def _config_json(args, c, world):
    """Names the workload; identical for both arms (the reference arm times a bounded sample of it)."""
    return {"workload": c["name"], "config": args.config, "n_sv": int(c["n_sv"]) * (world if args.weak else 1),
            "scaling_mode": "weak: n_sv per GPU" if args.weak else "strong: one SV list, LPT-partitioned by vapor_b200.multi.partition_svs",
            "reads_per_sv": READS_PER_SV, "seed": c["seed"],
            "l2": "inputs larger than L2 (sequence + k-mer word arrays >> 126 MB at the default sizes); no flush needed"}


# ------------------------------------------------------------------------------------------------
# CPU side: the reference implementation on host cores (test infrastructure; never on the GPU path)
# ------------------------------------------------------------------------------------------------
_CPU_IMPL = None


def _cpu_impl():
    """(module, kind): the compiled reference if oracle/_ref holds it, else the numpy oracle port."""
    global _CPU_IMPL
    if _CPU_IMPL is None:
        from oracle.reference_loader import load_reference
        mod = None
        try:
            mod = load_reference()
        except Exception:
            mod = None
        if mod is not None:
            _CPU_IMPL = (mod, "reference")
        else:
            from oracle import vapor_oracle
            _CPU_IMPL = (vapor_oracle, "port")
    return _CPU_IMPL


def _cpu_task(args):
    from oracle import batch_oracle as BO
    read, ref, alt, miss, k, mode = args
    impl, _ = _cpu_impl()
    try:
        sc, _ = BO.score_task(read, ref, alt, miss, k, mode, impl)
    except KeyError:
        sc = None
    return sc


def cpu_score_sample(batch, cores: int, pool=None):
    """Score every task of `batch` with the CPU implementation on `cores` processes; returns (seconds, scores)."""
    tasks = []
    for t in range(batch.n_task):
        tasks.append((batch.seq(int(batch.task_read[t])).decode("latin-1"), batch.seq(int(batch.task_ref[t])).decode("latin-1"),
                      batch.seq(int(batch.task_alt[t])).decode("latin-1"), int(batch.task_miss[t]), int(batch.task_k[t]),
                      int(batch.task_mode[t])))
    # longest first: the per-read cost spans two orders of magnitude
    order = sorted(range(len(tasks)), key=lambda i: -len(tasks[i][0]) * (len(tasks[i][1]) + len(tasks[i][2])))
    t0 = time.perf_counter()
    if cores <= 1 or pool is None:
        out = [_cpu_task(tasks[i]) for i in order]
    else:
        out = pool.map(_cpu_task, [tasks[i] for i in order], chunksize=1)
    dt = time.perf_counter() - t0
    scores = [None] * len(tasks)
    for i, o in zip(order, out):
        scores[i] = o
    # per-SV summary belongs to the path too (result_organize_ins + gt_estimate_log_likelihood)
    from oracle import vapor_oracle as O
    t1 = time.perf_counter()
    for s in range(batch.n_sv):
        a, b = int(batch.sv_task_off[s]), int(batch.sv_task_off[s + 1])
        O.summarize_sv([v for v in scores[a:b] if v is not None])
    dt += time.perf_counter() - t1
    return dt, scores


def _make_pool(cores):
    import multiprocessing as mp
    _cpu_impl()                                    # load before fork so workers inherit it
    return mp.get_context("fork").Pool(cores) if cores > 1 else None


def stratified_ids(costs: np.ndarray, n_pick: int, offset: int = 0) -> np.ndarray:
    """`n_pick` SVs spread evenly over the list ordered by cost (every cost decile and, because types cycle with
    the index, every SV type is represented); `offset` rotates the choice so successive samples are disjoint."""
    order = np.argsort(costs, kind="stable")
    n = len(order)
    n_pick = min(n_pick, n)
    pos = (np.arange(n_pick) * n // n_pick + offset) % n
    return np.sort(order[pos])


def _composition(w):
    types = sorted(set(w.sv_type))
    bins = [0, 500, 2000, 5000, 20000, 50000, 10 ** 9]
    out = {}
    for t in types:
        lens = w.sv_len[[i for i, x in enumerate(w.sv_type) if x == t]]
        hist = np.histogram(lens, bins=bins)[0]
        out[t] = {f"<{b}": int(h) for b, h in zip(bins[1:], hist) if h}
    return out


def run_reference_arm(args):
    """--impl reference: time the reference's CPU implementation on all host cores (rank 0 only).
    Every step scores a fresh stratified sample of the configured SV list, sized for a few seconds of all-core work."""
    rank, _, world = _dist_env()
    if rank != 0:
        return
    from vapor_b200 import synth
    c = _cfg(args)
    cores = os.cpu_count() or 1
    impl, kind = _cpu_impl()
    gk = _gen_kwargs(c)
    n_list = c["n_sv"] * (world if args.weak else 1)
    costs = synth.workload_costs(min(n_list, 20000), **{k: v for k, v in gk.items()})
    # cells one step should hold: about ref_seconds of all-core work at the reference's measured ~2.5e8 cells/s/core
    budget_cells = args.ref_seconds * cores * 2.5e8
    mean_cost = float(costs.mean())
    n_sample = args.ref_svs if args.ref_svs > 0 else int(max(4, min(len(costs) // max(1, args.warmup + args.steps), budget_cells / mean_cost)))
    pool = _make_pool(cores)
    times, reads, cells, n_svs = [], 0, 0, 0
    comp = {}
    single = None
    for it in range(args.warmup + args.steps):
        ids = stratified_ids(costs, n_sample, offset=it)
        w = synth.make_workload(0, sv_ids=ids, workers=min(cores, 16), **gk)
        if single is None:                                   # single-core rate on a small sub-sample (not part of the timed steps)
            sub = w.batch.shard(range(min(3, w.batch.n_sv)))
            dt1, _ = cpu_score_sample(sub, 1, None)
            single = {"reads_per_s": sub.n_task / dt1, "reads": int(sub.n_task)}
        dt, _ = cpu_score_sample(w.batch, cores, pool)
        if it >= args.warmup:
            times.append(dt); reads += w.batch.n_task; cells += w.cells; n_svs += w.batch.n_sv
            for t, d in _composition(w).items():
                for b, h in d.items():
                    comp.setdefault(t, {}).setdefault(b, 0)
                    comp[t][b] += h
    if pool is not None:
        pool.close()
    total = sum(times)
    value = reads / total
    sample = (f"{n_svs} SVs ({reads} reads, {cells:.3e} cells) in {len(times)} disjoint stratified samples of {n_sample} SVs of the "
              f"same seeded SV list (evenly spaced over the list ordered by cost), {total:.1f} s of wall time, "
              f"{getattr(impl, '__vapor_kind__', 'numpy-oracle')}, per-read tasks over a {cores}-process pool")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": _config_json(args, c, world),
        "sample": {"n_sv": n_svs, "reads": int(reads), "cells": int(cells), "wall_s": total, "composition_type_x_svlen": comp},
        "cells_per_sec": cells / total,
        "single_core": single,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            if t0 <= ts <= t1 + 0.2:
                try:
                    sm.append(float(f[0])); smax = float(f[1])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            for ts, row in self.rows[-3:]:
                f = [x.strip() for x in row.split(",")]
                try:
                    sm.append(float(f[0])); smax = float(f[1])
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def output_checksum(res) -> str:
    """SHA-256 over the gathered results that must not depend on how the list was sharded."""
    h = hashlib.sha256()
    for f in ("task_hitsum", "task_hits", "task_status", "task_score", "task_stat", "sv_gt", "sv_nscore", "sv_qs", "sv_gs", "sv_gq"):
        h.update(np.ascontiguousarray(getattr(res, f)).tobytes())
    return h.hexdigest()


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    rank, local_rank, world = _dist_env()
    # the synthetic workload is made first: it forks worker processes, which is safest before CUDA / NCCL exist
    from vapor_b200 import multi, synth
    from vapor_b200.engine import Engine, Pipeline
    c = _cfg(args)
    gk = _gen_kwargs(c)
    cores = os.cpu_count() or 1
    t_gen = time.perf_counter()
    if args.weak:
        n_list = c["n_sv"] * world
        my_ids = np.arange(rank * c["n_sv"], (rank + 1) * c["n_sv"], dtype=np.int64)
        costs, ntask = synth.workload_costs(n_list, with_tasks=True, **gk) if world > 1 else (None, None)
        parts = [np.arange(r * c["n_sv"], (r + 1) * c["n_sv"], dtype=np.int64) for r in range(world)]
    else:
        n_list = c["n_sv"]
        costs, ntask = synth.workload_costs(n_list, with_tasks=True, **gk)
        parts = multi.partition_svs(costs, world)            # every rank computes the same partition
        my_ids = parts[rank]
    w = synth.make_workload(0, sv_ids=my_ids, workers=max(1, min(32, cores // max(world, 1))), **gk)
    t_gen = time.perf_counter() - t_gen
    if costs is not None:
        part_cells = np.array([float(costs[p].sum()) for p in parts])
        sv_task_off = np.zeros(n_list + 1, dtype=np.int64)
        np.cumsum(ntask, out=sv_task_off[1:])
    else:
        part_cells = np.array([float(w.cells)])
        sv_task_off = w.batch.sv_task_off
    n_task_list = int(sv_task_off[-1])

    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize(local_rank)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

    def reduce(x: float, op) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce(x, dist.ReduceOp.MAX) if dist is not None else x

    def sum_over_ranks(x):
        return reduce(x, dist.ReduceOp.SUM) if dist is not None else x

    numa_node = None
    if world > 1 and not args.no_numa_bind:
        # one process per GPU: pinned buffers and planning threads on the NUMA node the GPU hangs off (no-op when the host
        # shows no topology); the CPU baseline runs only at N = 1 and is never restricted
        from vapor_b200.engine import bind_to_gpu_numa
        numa_node = bind_to_gpu_numa(local_rank)
    eng = Engine(local_rank)
    opts = {}
    if args.tile_variant >= 0:
        opts["tile_variant"] = args.tile_variant
    if args.k2_mode >= 0:
        opts["k2_mode"] = args.k2_mode
    if args.k3_mode >= 0:
        opts["k3_mode"] = args.k3_mode
    for kv in args.opt:
        k_, v_ = kv.split("=")
        opts[k_] = int(v_)
    if args.hit_budget_gb > 0:
        opts["hit_budget_bytes"] = int(args.hit_budget_gb * (1 << 30))
    if args.k2_ctas_per_sm > 0:
        opts["k2_ctas_per_sm"] = args.k2_ctas_per_sm
    for k_, v_ in opts.items():
        eng.set_option(k_, v_)
    batch = eng.pin_batch(w.batch)                      # inputs live in pinned host memory
    res = eng.pinned_results(batch.n_task, batch.n_sv)
    n_reads = batch.n_task

    # the gathered results of the whole list, in input order, shared by all ranks (rank 0 owns them)
    tag = f"{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}"
    shared = None
    if world > 1:
        if rank == 0:
            shared = multi.SharedResults(tag, n_task_list, n_list, create=True)
        barrier()
        if rank != 0:
            shared = multi.SharedResults(tag, n_task_list, n_list, create=False)
        my_tix = multi.part_task_index(sv_task_off, my_ids)

    def gather(r):
        """This rank's results to their input positions in the shared arrays (at N = 1 they are in input order already)."""
        if shared is not None:
            multi.scatter_part(shared.results, my_tix, my_ids, r, sv_task_off=sv_task_off)

    # ---- resident timing: upload once, K x run() ---------------------------------------------------
    eng.upload(batch)
    for _ in range(args.warmup):
        eng.run()
    barrier()
    sampler = ClockSampler(local_rank)
    t0 = time.perf_counter()
    acc = {"total_ms": 0.0, "tile_ms": 0.0, "pack_ms": 0.0, "table_ms": 0.0, "score_ms": 0.0, "score_warp_ms": 0.0, "genotype_ms": 0.0, "launches": 0}
    for _ in range(args.steps):
        eng.run()
        tm = eng.timings()
        for k_ in acc:
            acc[k_] += tm[k_]
    barrier()
    t1 = time.perf_counter()
    wall_resident = t1 - t0
    clocks = sampler.stop(t0, t1)
    eng.fetch(res)
    tm_last = eng.timings()
    dev_s = max_over_ranks(acc["total_ms"] * 1e-3)
    total_reads = sum_over_ranks(float(n_reads))
    total_cells = sum_over_ranks(float(tm_last["cells"]))
    total_eval = sum_over_ranks(float(tm_last["evaluated_cells"]))
    value = total_reads * args.steps / dev_s

    # ---- end to end: host buffers -> public API -> host results in input order, every step -----------------------
    # (a) one blocking call per step (Engine.score_into = vapor_gpu_score): plan + H2D + kernels + D2H, then the gather
    for _ in range(max(1, args.warmup // 2)):
        gather(eng.score_into(batch, res))
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        gather(eng.score_into(batch, res))
    barrier()
    e2e_single_s = max_over_ranks(time.perf_counter() - te0)
    tm_e2e = eng.timings()
    # (b) the same steps through engine.Pipeline (two handles, double-buffered): every step still plans, copies its
    #     inputs in from pinned host memory, copies its results out and gathers them, overlapped with the previous step's kernels
    pipe = Pipeline(local_rank, depth=2, options=opts)
    res2 = [res, eng.pinned_results(batch.n_task, batch.n_sv)]
    pipe.map([batch] * 2, res2, after=gather)                           # warm-up: buffers of both handles allocated
    barrier()
    te0 = time.perf_counter()
    pipe.map([batch] * args.steps, [res2[i % 2] for i in range(args.steps)], after=gather)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - te0)
    stage_ms = {k_: round(1e3 * v_ / args.steps, 3) for k_, v_ in pipe.stage_s.items()}      # rank 0's, per step
    pipe.close()
    e2e_value = total_reads * args.steps / e2e_s
    launches_total = sum_over_ranks(float(acc["launches"]))
    gathered = shared.results if shared is not None else res
    checksum = output_checksum(gathered) if rank == 0 else None

    # ---- roofline ---------------------------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    K = args.steps
    n_waves = max(1, int(tm_last["n_waves"]))
    mode = int(tm_last["k2_mode"])
    nominal = 148 * 128 * 1.965e9                              # 4 SMSPs x 1 warp instruction/clk x 32 lanes at 1965 MHz
    kernels = {}
    # kernel 1: 1 B/base read + 0.25 B/base of 2-bit code = SURVEY 8(d)'s algorithmic 1.25 B/base; what it really moves is
    # 1 B read + 1 B code + 4 B word per base
    pack_s = acc["pack_ms"] * 1e-3 / K
    if pack_s > 0:
        kernels["k1_pack_kmers"] = {"bound": "hbm", "ms": 1e3 * pack_s, "algorithmic_bytes": 1.25 * tm_last["bases"], "actual_bytes": 6.0 * tm_last["bases"],
                                    "achieved": 1.25 * tm_last["bases"] / pack_s / 1e9, "achieved_actual_bytes": 6.0 * tm_last["bases"] / pack_s / 1e9,
                                    "peak": hbm_peak, "unit": "GB/s", "frac": 1.25 * tm_last["bases"] / pack_s / 1e9 / hbm_peak,
                                    "frac_actual_bytes": 6.0 * tm_last["bases"] / pack_s / 1e9 / hbm_peak}
    tile_s = acc["tile_ms"] * 1e-3 / K
    k_arr = batch.task_k.astype(np.int64)
    if mode == 0:
        ops_per_cell = float(np.mean((2 * k_arr + 31) // 32)) if int(k_arr.max(initial=10)) <= 15 else 1.0
        int_peak = max(eng.int_peak(3), eng.int_peak(6), eng.int_peak(1), eng.int_peak(2))
        ach = tm_last["cells"] * ops_per_cell / tile_s
        kernels["k2_tile_match"] = {"bound": "int32_issue", "ms": 1e3 * tile_s, "achieved": ach / 1e9, "peak": nominal / 1e9, "unit": "Gop/s",
                                    "frac": ach / nominal, "peak_source": "nominal issue limit 148 SM x 128 lanes x 1.965 GHz",
                                    "frac_of_measured_dual_pipe_streams": ach / int_peak, "measured_dual_pipe_gops": int_peak / 1e9,
                                    "measured_peak_is": "best of independent LOP3+IMAD / ISETP+IMAD streams (vapor_gpu_int_peak 3, 6) -- not the kernel's own loop",
                                    "ops": "one 32-bit word compare per cell (k <= 15: exact canonical word; k > 15: hashed word, confirmed only on a match)",
                                    "padded_cells": int(tm_last["padded_cells"]), "launches_per_step": n_waves}
    else:
        # join kernel: per launch it must read every plot's read words once (4 B), every table once and write the hits (8 B)
        n_words = float(tm_last["probe_words"])
        table_bytes = float(tm_last["table_bytes"])
        alg = 4.0 * n_words + table_bytes + 8.0 * tm_last["hits"]
        kernels["k2_join_match"] = {"bound": "hbm", "ms": 1e3 * tile_s, "algorithmic_bytes": alg, "achieved": alg / tile_s / 1e9, "peak": hbm_peak,
                                    "unit": "GB/s", "frac": alg / tile_s / 1e9 / hbm_peak,
                                    "cells_nominal": int(tm_last["cells"]), "cells_evaluated": int(tm_last["evaluated_cells"]),
                                    "evaluated_cells_per_s": tm_last["evaluated_cells"] / tile_s,
                                    "read_words_probed_per_s": n_words / tile_s,
                                    "note": "radix-partitioned join: a read k-mer first meets the table's membership bitmap; the survivors (one in three) "
                                            "are compacted into full-warp rounds, each against the bucket it falls into; bound by global-load latency and "
                                            "instruction issue (ncu: issue slots 51-54 % busy), not by HBM or the compare count"}
        table_s = tm_last["table_ms"] * 1e-3
        if table_s > 0:
            alg1b = table_bytes * 10.0 / 6.0            # 4 B/word read + 6 B/word (+ offsets) written
            kernels["k1b_build_tables"] = {"bound": "hbm", "ms": 1e3 * table_s, "algorithmic_bytes": alg1b,
                                           "achieved": alg1b / table_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                           "frac": alg1b / table_s / 1e9 / hbm_peak}
    # kernel 3 is two kernels: the warp-per-task one (value ranges up to 8 192 bins by default) and the CTA-per-task one; the library
    # times them apart (score_warp_ms).  Which kernel a task goes to follows from its plots' value range n + m - 1 (api.cu,
    # k3_class_of), so the hits each kernel must read are known exactly from the fetched per-task hit counts.
    score_s = acc["score_ms"] * 1e-3 / K
    score_w_s = acc["score_warp_ms"] * 1e-3 / K
    if score_s > 0:
        seq_len = np.diff(batch.seq_off)
        kk = batch.task_k.astype(np.int64)
        n_read = np.maximum(0, seq_len[batch.task_read] - kk + 1)
        miss = batch.task_miss.astype(np.int64)
        nb_task = np.ones(batch.n_task, dtype=np.int64)
        for which in (batch.task_ref, batch.task_alt):
            ls = seq_len[which]
            cut = np.where(miss >= 0, miss, np.maximum(0, ls + miss))
            m_struct = np.maximum(0, ls - kk + 1 - cut)
            nb_task = np.maximum(nb_task, n_read + m_struct - 1)
        warp_cap = int(opts.get("k3_warp_classes", 3))
        warp_cap = 0 if int(opts.get("k3_mode", 1)) == 0 or warp_cap == 0 else (2048, 4096, 8192, 16384, 26624)[warp_cap - 1]
        on_warp = nb_task <= warp_cap
        # every distinct plot's hits once: columns 2-3 repeat columns 0-1 when the W10 opinion looks at the same plots
        hits_task = res.task_hits[:, :2].astype(np.int64).sum(axis=1)
        note3 = ("must read every hit once (8 B) and write 77 B per read; really 4-10 passes over a plot's hits (L2-resident) with short "
                 "dependent phases: bound by latency and instruction issue")
        for name, sel, sec in (("k3w_score_reads", on_warp, score_w_s), ("k3_score_reads", ~on_warp, score_s - score_w_s)):
            if sec <= 0 or not sel.any():
                continue
            algk = 8.0 * float(hits_task[sel].sum()) + 77.0 * float(sel.sum())
            kernels[name] = {"bound": "hbm", "ms": 1e3 * sec, "algorithmic_bytes": algk, "achieved": algk / sec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": algk / sec / 1e9 / hbm_peak, "tasks": int(sel.sum()), "hits": int(hits_task[sel].sum()), "note": note3}
    dom = max(kernels, key=lambda n: kernels[n]["ms"])
    step_ms = acc["total_ms"] / K
    roofline = dict(kernels[dom])
    # DRAM traffic of the dominant kernel, measured by ncu --set full (dram__bytes_read + dram__bytes_write), per launch like
    # `achieved`.  Two captures exist: one WAVE of the default workload (a wave of the 100k-SV list = 6 250 SVs; the wave's
    # launches of the kernel averaged) -- quoted as `traffic` when this run consists of such waves -- and a whole run of 2 000 SVs
    # of config 2, quoted only for that very workload.  Nothing is scaled.
    traffic, traffic_profile = None, None
    n_launch_dom = {"k1_pack_kmers": 1, "k1b_build_tables": 1, "k2_join_match": 3, "k2_tile_match": 1, "k3_score_reads": 2, "k3w_score_reads": 3}.get(dom, 1) * n_waves
    roofline["launches_per_step"] = n_launch_dom
    roofline["algorithmic_bytes_per_launch"] = roofline.get("algorithmic_bytes", 0.0) / max(1, n_launch_dom) if "algorithmic_bytes" in roofline else None
    roofline["ms_per_launch"] = roofline["ms"] / max(1, n_launch_dom)

    def _kernel_traffic(tp):
        return tp["kernels"].get(dom)
    try:
        tw = json.load(open(os.path.join(ROOT, "profiles", "r02h_wave_traffic.json")))
        kt = _kernel_traffic(tw)
        if kt and args.config == tw["config"] and not args.weak and w.batch.n_sv >= tw["wave_svs"] and mode == 1:
            traffic = (kt["dram_bytes_read"] + kt["dram_bytes_write"]) / max(1, kt["launches"])
            traffic_profile = {"measured_on": tw["source"], "dram_bytes_per_wave": kt["dram_bytes_read"] + kt["dram_bytes_write"],
                               "launches_per_wave": kt["launches"], "kernel_ms_under_ncu_per_wave": kt["ms"]}
    except Exception:
        pass
    try:
        tp = json.load(open(os.path.join(ROOT, "profiles", "r02g_traffic.json")))
        kt = _kernel_traffic(tp)
        if kt and traffic_profile is None:
            traffic_profile = {"measured_on": f"config {tp['config']}, {tp['n_sv']} SVs (ncu --set full)", "dram_bytes": kt["dram_bytes_read"] + kt["dram_bytes_write"],
                               "launches": kt["launches"], "kernel_ms_under_ncu": kt["ms"]}
            if args.config == tp["config"] and n_list == tp["n_sv"] and world == 1:
                traffic = (kt["dram_bytes_read"] + kt["dram_bytes_write"]) / max(1, kt["launches"])
    except Exception:
        pass
    roofline.update({"kernel": dom, "share_of_step": kernels[dom]["ms"] / max(step_ms, 1e-9), "peak_source": kernels[dom].get("peak_source", hbm_src),
                     "traffic": traffic, "traffic_profile": traffic_profile})
    if roofline["bound"] != "hbm":
        roofline["bound_schema"] = "integer issue (the schema's hbm|tensor does not apply: integer compare work)"

    # ---- CPU baseline on rank 0 at N=1 ----------------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        impl, kind = _cpu_impl()
        budget_cells = args.ref_seconds * 3 * cores * 2.5e8
        my_costs = multi.sv_costs(w.batch)
        n_sample = args.ref_svs if args.ref_svs > 0 else int(max(4, min(w.batch.n_sv, budget_cells / max(1.0, float(my_costs.mean())))))
        ids = stratified_ids(my_costs, n_sample)
        sub = w.batch.shard(ids)
        pool = _make_pool(cores)
        dt, cpu_scores = cpu_score_sample(sub, cores, pool)
        if pool is not None:
            pool.close()
        # the sample doubles as a parity spot check of the timed GPU results
        tix = multi.part_task_index(w.batch.sv_task_off, ids)
        ok = True
        for t, sc in zip(tix, cpu_scores):
            g_ok = res.task_status[t] == 1
            if (sc is None) != (not g_ok) or (sc is not None and abs(sc - res.task_score[t]) > 1e-5):
                ok = False
        cpu_baseline = {"value": sub.n_task / dt, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{n_sample} SVs ({sub.n_task} reads) spread evenly over the benchmark SV list ordered by cost, one pass "
                                  f"({dt:.1f} s), per-read tasks over a {cores}-process pool, {getattr(impl, '__vapor_kind__', 'numpy-oracle')}",
                        "gpu_scores_match_on_sample": ok}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": _config_json(args, c, world),
            "workload_stats": {"svs": int(n_list), "reads": int(total_reads), "cells": int(total_cells),
                               "sequence_bytes_rank0": int(w.batch.seq_bytes.nbytes)},
            "cells_per_sec_nominal": total_cells * args.steps / dev_s,
            "cells_evaluated_per_sec": total_eval * args.steps / dev_s,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(batch.h2d_bytes()),
                    "d2h_bytes_per_step": int(res.d2h_bytes()), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "vapor_b200.engine.Pipeline(depth=2).map: two handles, each step = vapor_gpu_upload + vapor_gpu_run + vapor_gpu_fetch on "
                           "pinned host buffers (host planning + H2D + kernels + D2H every step, the steps of the two handles interleaved), "
                           "then multi.scatter_part into the shared input-order result arrays (N > 1)",
                    "bytes_are": "per rank (rank 0)", "pipeline_stage_ms_rank0": stage_ms,
                    "single_blocking_call": {"value": total_reads * args.steps / e2e_single_s, "ms_per_step": 1e3 * e2e_single_s / args.steps,
                                             "host_prep_ms": tm_e2e["host_prep_ms"], "h2d_ms": tm_e2e["h2d_ms"], "d2h_ms": tm_e2e["d2h_ms"]}},
            "gpu_launches": int(launches_total),
            "output_checksum": checksum,
            "partition": {"parts": world, "cells_max_over_mean": float(part_cells.max() / max(part_cells.mean(), 1.0)),
                          "rule": "identity" if world == 1 else ("contiguous blocks (--weak)" if args.weak else "multi.partition_svs: greedy LPT on cells per SV")},
            "roofline": roofline, "kernels": kernels,
            "cpu_baseline": cpu_baseline,
            "phase_ms_per_step": {"pack": acc["pack_ms"] / K, "table": acc["table_ms"] / K, "tile": acc["tile_ms"] / K,
                                  "score": acc["score_ms"] / K, "genotype": acc["genotype_ms"] / K},
            "wall_ms_per_step_resident": 1e3 * wall_resident / args.steps,
            "hits_per_step_rank0": int(tm_last["hits"]), "workload_gen_s": t_gen, "numa_node_rank0": numa_node,
            "k2_mode": mode, "n_waves_rank0": n_waves, "overflow_plots_rank0": int(tm_last["n_overflow_plots"]),
            "sv_called": {"gt_0/0": int((gathered.sv_gt == 0).sum()), "gt_0/1": int((gathered.sv_gt == 1).sum()),
                          "gt_1/1": int((gathered.sv_gt == 2).sum()), "NA": int((gathered.sv_gt == 255).sum())},
        }
        _emit(line)
    eng.close()
    if dist is not None:
        dist.barrier()
    if shared is not None:
        shared.close()
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _emit(line: dict):
    """The one JSON line goes to the real stdout; everything else any library prints (NCCL banners ...) was sent to stderr."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                                   # stdout of this process and its children -> stderr from here on
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=sorted(CONFIGS), help="BASELINE.json workload (see the module docstring)")
    ap.add_argument("--n-sv", type=int, default=0, help="SVs in the list (0 = the config's size); per GPU with --weak")
    ap.add_argument("--weak", action="store_true", help="every rank scores its own --n-sv SVs instead of a share of one list")
    ap.add_argument("--ref-svs", type=int, default=0, help="SVs per CPU sample (0 = sized from --ref-seconds)")
    ap.add_argument("--ref-seconds", type=float, default=10.0, help="all-core seconds one reference-arm step should take")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tile-variant", type=int, default=-1, help="tile-kernel inner loop (-1 = library default)")
    ap.add_argument("--k2-mode", type=int, default=-1, help="kernel 2: 1 = join (library default), 0 = all-pairs tile kernel")
    ap.add_argument("--k3-mode", type=int, default=-1, help="kernel 3: 1 = warp per task (library default), 0 = CTA per task everywhere")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not bind each rank to its GPU's NUMA node")
    ap.add_argument("--opt", action="append", default=[], help="library tunable name=value (vapor_gpu_set_option), repeatable")
    ap.add_argument("--hit-budget-gb", type=float, default=0, help="device memory for the hit slab of one wave (0 = library default)")
    ap.add_argument("--k2-ctas-per-sm", type=int, default=0, help="persistent-grid size of the tile kernel (0 = occupancy)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch ourselves one rank per GPU (the driver already does this with torchrun)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=_REAL_STDOUT.fileno()))
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
