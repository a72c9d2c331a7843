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
