"""Host-side logic: batch packing, sharding, synthetic workload determinism, multi-GPU merge (gloo)."""
import os
import sys

import numpy as np
import pytest

from oracle import batch_oracle as BO
from vapor_b200 import multi, synth
from vapor_b200.engine import Batch, MODE_ABS, MODE_W10, Results, _alloc_results


def test_batch_pack_roundtrip():
    b = Batch()
    r0 = b.add_seq("ACGTACGTAC"); s0 = b.add_seq(b"ACGTTTTT"); s1 = b.add_seq(np.frombuffer(b"GGGG", np.uint8))
    b.add_task(r0, s0, s1, 2, 10, MODE_ABS)
    b.end_sv("sv0")
    b.end_sv("sv1-empty")
    b.add_task(r0, s1, s0, 0, 20, MODE_W10)
    pb = b.pack()
    assert pb.n_seq == 3 and pb.n_task == 2 and pb.n_sv == 3
    assert pb.seq(0) == b"ACGTACGTAC" and pb.seq(2) == b"GGGG"
    assert pb.sv_task_off.tolist() == [0, 1, 1, 2]
    assert pb.task_k.tolist() == [10, 20] and pb.task_miss.tolist() == [2, 0]
    c = pb.c_struct()
    assert c.n_task == 2 and c.n_sv == 3


def test_workload_is_deterministic_and_prefix_stable():
    a = synth.make_workload(6, seed=5, size_range=(50, 400), reads_per_sv=3)
    b = synth.make_workload(6, seed=5, size_range=(50, 400), reads_per_sv=3, workers=2)
    assert np.array_equal(a.batch.seq_bytes, b.batch.seq_bytes)
    c = synth.make_workload(2, seed=5, size_range=(50, 400), reads_per_sv=3, first_sv=3)
    sub = a.batch.shard([3, 4])
    assert np.array_equal(sub.seq_bytes, c.batch.seq_bytes) and np.array_equal(sub.task_read, c.batch.task_read)
    assert a.cells == int(multi.sv_costs(a.batch).sum())


def test_partition_balanced_and_ordered():
    costs = [100, 1, 1, 50, 50, 1, 99, 2]
    parts = multi.partition_svs(costs, 3)
    assert sorted(np.concatenate(parts).tolist()) == list(range(8))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) <= 104
    for p in parts:
        assert p.tolist() == sorted(p.tolist())


def _oracle_scorer(pb):
    e = BO.score_batch(pb, with_hits=False)
    r = _alloc_results(pb.n_task, pb.n_sv)
    for f in r.__dataclass_fields__:
        getattr(r, f)[...] = e[f]
    return r


def test_sharded_scoring_keeps_input_order():
    w = synth.make_workload(7, seed=9, size_range=(50, 300), reads_per_sv=3)
    whole = _oracle_scorer(w.batch)
    merged = multi.score_sharded(w.batch, [_oracle_scorer] * 3)
    for f in whole.__dataclass_fields__:
        assert np.array_equal(getattr(whole, f), getattr(merged, f)), f


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.make_workload(6, seed=13, size_range=(50, 300), reads_per_sv=3)
    out = multi.score_distributed(w.batch, _oracle_scorer, rank, world, dist)
    if rank == 0:
        q.put({f: getattr(out, f) for f in out.__dataclass_fields__})
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_world2_gloo():
    """world_size 2 over gloo on CPU: SV-sharded scoring gathers to rank 0 in input order.  (The scorer is
    the CPU oracle here -- this tests the sharding/gather plumbing, not the kernels.)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = synth.make_workload(6, seed=13, size_range=(50, 300), reads_per_sv=3)
    whole = _oracle_scorer(w.batch)
    for f in whole.__dataclass_fields__:
        assert np.array_equal(getattr(whole, f), got[f]), f


def _strong_worker(rank, world, port, tag, q):
    """What bench.py does per rank under torchrun (strong scaling), with the CPU oracle standing in for the GPU."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = dict(seed=31, recipe="complex", size_range=(200, 700), reads_per_sv=3)
    n_list = 9
    costs, ntask = synth.workload_costs(n_list, with_tasks=True, **kw)
    parts = multi.partition_svs(costs, world)
    mine = parts[rank]
    w = synth.make_workload(0, sv_ids=mine, **kw)
    sv_task_off = np.concatenate([[0], np.cumsum(ntask)])
    shared = multi.SharedResults(tag, int(sv_task_off[-1]), n_list, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        shared = multi.SharedResults(tag, int(sv_task_off[-1]), n_list, create=False)
    part = _oracle_scorer(w.batch)
    multi.scatter_part(shared.results, multi.part_task_index(sv_task_off, mine), mine, part)                   # numpy route
    check = _alloc_results(int(sv_task_off[-1]), n_list)
    multi.scatter_part(check, None, mine, part, sv_task_off=sv_task_off)                                        # native route
    tix = multi.part_task_index(sv_task_off, mine)
    for f in part.__dataclass_fields__:
        sel = tix if f.startswith("task_") else mine
        assert np.array_equal(getattr(check, f)[sel], getattr(shared.results, f)[sel]), f
    multi.scatter_part(shared.results, None, mine, part, sv_task_off=sv_task_off)
    dist.barrier()
    if rank == 0:
        q.put({f: np.array(getattr(shared.results, f)) for f in shared.results.__dataclass_fields__})
    dist.barrier()
    shared.close()
    dist.destroy_process_group()


def test_strong_scaling_world2_shared_gather():
    """world_size 2 over gloo on CPU: every rank derives the same LPT partition from costs that need no sequence,
    builds and scores only its part, and writes it to its input positions in /dev/shm arrays; rank 0 then holds the
    whole list in input order -- identical (same bench.output_checksum) to scoring the list in one piece."""
    import torch.multiprocessing as mp
    import bench
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + os.getpid() % 300
    tag = f"test_{os.getpid()}"
    procs = [ctx.Process(target=_strong_worker, args=(r, 2, port, tag, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole_w = synth.make_workload(9, seed=31, recipe="complex", size_range=(200, 700), reads_per_sv=3)
    whole = _oracle_scorer(whole_w.batch)
    for f in whole.__dataclass_fields__:
        assert np.array_equal(getattr(whole, f), got[f]), f
    assert bench.output_checksum(whole) == bench.output_checksum(Results(**got))
    assert not [f for f in os.listdir("/dev/shm") if tag in f]                  # the owner removed the shared arrays


def test_costs_without_sequences_match_generated_workloads():
    """synth.workload_costs (used to partition a 100k-SV list before any rank builds its share) equals the cells of
    the generated tasks, for every recipe; an arbitrary subset of SV ids regenerates exactly those SVs."""
    for recipe, kw in (("simple", dict(size_range=(50, 900))), ("complex", dict(size_range=(200, 900))),
                       ("large", dict(size_range=(10000, 14000), k_choices=(10, 20, 30, 40)))):
        n = 14 if recipe == "complex" else 6
        w = synth.make_workload(n, seed=3, recipe=recipe, reads_per_sv=2, **kw)
        costs, ntask = synth.workload_costs(n, 3, recipe=recipe, reads_per_sv=2, with_tasks=True, **kw)
        assert np.array_equal(costs, multi.sv_costs(w.batch)), recipe
        assert np.array_equal(ntask, np.diff(w.batch.sv_task_off)), recipe
        ids = [2, 3, n - 1]
        sub = synth.make_workload(0, seed=3, recipe=recipe, reads_per_sv=2, sv_ids=ids, **kw)
        ref = w.batch.shard(ids)
        assert np.array_equal(sub.batch.seq_bytes, ref.seq_bytes) and np.array_equal(sub.batch.task_alt, ref.task_alt), recipe


def test_complex_recipe_shapes():
    """Config 3 events: one task group per distinct alternative haplotype, the same reads against each; REDEF exactly
    when a block repeats in the allele (Simple_function.pyx:1519-1523); junction fallbacks are 1 kb W10 windows."""
    from vapor_b200.engine import MODE_REDEF
    w = synth.make_workload(14, seed=11, recipe="complex", size_range=(300, 900), reads_per_sv=4)
    off = w.batch.sv_task_off
    lens = np.diff(w.batch.seq_off)
    for s, kind in enumerate(w.sv_type):
        alleles = synth.COMPLEX_KINDS[kind][1]
        t = slice(off[s], off[s + 1])
        assert off[s + 1] - off[s] == 4 * len(alleles), kind
        assert len(set(w.batch.task_ref[t].tolist())) == 1 and len(set(w.batch.task_alt[t].tolist())) == len(alleles)
        for a, al in enumerate(alleles):
            ta = slice(off[s] + 4 * a, off[s] + 4 * a + 4)
            assert np.array_equal(w.batch.task_read[ta], w.batch.task_read[off[s]:off[s] + 4])       # same reads for every allele
            if kind == "JUNCTION":
                assert set(w.batch.task_mode[ta].tolist()) == {MODE_W10} and lens[w.batch.task_ref[off[s]]] == 1001
            else:
                repeated = max(al.count(c) for c in al if c != "^") > 1
                assert set(w.batch.task_mode[ta].tolist()) == {MODE_REDEF if repeated else MODE_ABS}, (kind, al)
    # reads drawn from the planted haplotypes score: most het / hom-alt events get a non-0/0 call from the oracle
    exp = BO.score_batch(w.batch, with_hits=False)
    called = [(g, int(c)) for g, c in zip(w.sv_genotype, exp["sv_gt"]) if g > 0]
    assert sum(1 for g, c in called if c in (1, 2)) >= 0.7 * len(called)


def test_stratified_sample_covers_types_and_sizes():
    import bench
    costs = synth.workload_costs(400, 5)
    ids = bench.stratified_ids(costs, 40)
    assert len(set(ids.tolist())) == 40
    assert len({i % 4 for i in ids}) == 4                                          # all four SV types
    q = np.quantile(costs, [0.25, 0.75])
    assert (costs[ids] < q[0]).any() and (costs[ids] > q[1]).any()
    assert not set(ids.tolist()) & set(bench.stratified_ids(costs, 40, offset=1).tolist())


def test_bam_in_decide_patterns(tmp_path):
    """One file, or every file of the directory matching an XXX / * pattern (Simple_function.pyx:69-89)."""
    from vapor_b200 import Simple_function as SF
    for name in ("s.chr1.bam", "s.chr2.bam", "s.chr1.bai", "other.chr1.bam"):
        (tmp_path / name).write_text("x")
    one = str(tmp_path / "s.chr1.bam")
    assert SF.bam_in_decide(one, None) == [one]
    assert sorted(SF.bam_in_decide(str(tmp_path / "s.XXX.bam"), None)) == [str(tmp_path / "s.chr1.bam"), str(tmp_path / "s.chr2.bam")]
    assert sorted(SF.bam_in_decide(str(tmp_path / "s.*.bam"), None)) == [str(tmp_path / "s.chr1.bam"), str(tmp_path / "s.chr2.bam")]


def test_small_helpers_match_reference_semantics():
    from vapor_b200 import Simple_function as SF
    assert SF.complementary("ACGTNacgtnXRY") == "TGCANtgcan"            # quirk: other characters are dropped
    assert SF.reverse("ACG") == "GCA"
    assert SF.flank_length_calculate(["c", 100, 150]) == 50 and SF.flank_length_calculate(["c", 100, 5000]) == 500
    assert SF.letter_split("c^ba") == ["c^", "b", "a"]
    assert SF.block_subsplot(["chr1", "10", "20", "chr2", "5", "9"], ["chr1", "chr2"]) == [["chr1", 10, 20], ["chr2", 5, 9]]
    h = SF.bp_to_chr_hash(["chr1", 100, 200, 300], ["chr1"], 50)
    assert h["a"] == ["chr1", 100, 200] and h["b"] == ["chr1", 200, 300] and h["+"] == ["chr1", 300, "350"] and h["-"] == ["chr1", "50", 100]
    assert SF.block_around_check("ba", "ab") == [["-", "b"], ["b", "a"], ["a", "+"]]
    assert SF.minimize_pacbio_read_list([["r", i % 3, "q"] for i in range(30)])[:10] == [["r", 0, "q"]] * 10
    assert len(SF.minimize_pacbio_read_list([["r", i % 3, "q"] for i in range(30)])) == 20


def test_host_plan_independent_of_thread_count():
    """vapor_host_plan (no device needed): the plan digest -- operands, plots, tasks, waves, table chunks, join items,
    kernel-3 class lists -- is the same for 1, 2, 3 and 7 planning threads, for both kernel-2 variants, with one wave
    and with many."""
    from vapor_b200.engine import host_plan
    w = synth.make_workload(300, seed=41, size_range=(50, 1500), reads_per_sv=5, lowercase_every=7, max_miss=3)
    for mode in (0, 1):
        for budget in (0, 8 << 20):
            base = host_plan(w.batch, mode, 1, budget)
            assert base["tasks"] == w.batch.n_task and base["cells"] >= w.cells      # lower-case DEL structures add W10 plots
            assert (base["waves"] > 3) == (budget > 0)
            for th in (2, 3, 7):
                got = host_plan(w.batch, mode, th, budget)
                assert got["digest"] == base["digest"] and got["waves"] == base["waves"], (mode, budget, th)
    cx = synth.make_workload(28, seed=5, recipe="complex", size_range=(200, 900), reads_per_sv=3)
    assert host_plan(cx.batch, 1, 1)["digest"] == host_plan(cx.batch, 1, 4)["digest"]


def test_bind_to_gpu_numa_reads_sysfs_and_is_a_noop_without_topology(tmp_path, monkeypatch):
    """engine.bind_to_gpu_numa: the CPU list of the GPU's NUMA node from a (fake) sysfs tree; None, and no affinity
    change, when the bus id or the node is not visible."""
    import os
    from vapor_b200 import engine

    class FakeLib:
        def __init__(self, bus, rc=0): self.bus, self.rc = bus, rc
        def vapor_gpu_pci_bus_id(self, dev, buf, n):
            buf.value = self.bus.encode()
            return self.rc
    orig = os.sched_getaffinity(0)
    cpu = sorted(orig)[0]
    (tmp_path / "bus/pci/devices/0000:1b:00.0").mkdir(parents=True)
    (tmp_path / "bus/pci/devices/0000:1b:00.0/numa_node").write_text("1\n")
    (tmp_path / "devices/system/node/node1").mkdir(parents=True)
    (tmp_path / "devices/system/node/node1/cpulist").write_text(f"{cpu}-{cpu},9999\n")
    assert engine._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    try:
        monkeypatch.setattr(engine.N, "load", lambda: FakeLib("0000:1B:00.0"))
        assert engine.bind_to_gpu_numa(0, sysfs=str(tmp_path)) == 1
        assert os.sched_getaffinity(0) == {cpu}
        os.sched_setaffinity(0, orig)
        (tmp_path / "bus/pci/devices/0000:1b:00.0/numa_node").write_text("-1\n")
        assert engine.bind_to_gpu_numa(0, sysfs=str(tmp_path)) is None
        monkeypatch.setattr(engine.N, "load", lambda: FakeLib("0000:ff:00.0"))
        assert engine.bind_to_gpu_numa(0, sysfs=str(tmp_path)) is None
        monkeypatch.setattr(engine.N, "load", lambda: FakeLib("", rc=-1))
        assert engine.bind_to_gpu_numa(0, sysfs=str(tmp_path)) is None
        assert os.sched_getaffinity(0) == orig
    finally:
        os.sched_setaffinity(0, orig)
