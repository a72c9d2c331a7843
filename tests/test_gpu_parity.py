"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle on the
same seeded inputs.  Bars (BASELINE.json north_star): hit counts and coordinates bit-exact,
qs/gs/Rec within 1e-5 absolute, GT identical, GQ within 1e-3."""
import numpy as np
import pytest

from oracle import batch_oracle as BO
from oracle import vapor_oracle as O
from vapor_b200 import synth
from vapor_b200.engine import Batch, MODE_ABS, MODE_ABS_AND_W10, MODE_REDEF, MODE_W10

pytestmark = pytest.mark.gpu

QS_TOL = 1e-5     # north_star: VaPoR_qs/gs/Rec within 1e-5 absolute
GQ_TOL = 1e-3     # north_star: VaPoR_GQ within 1e-3


def _compare(res, exp, exact_scores=True):
    np.testing.assert_array_equal(res.task_hits, exp["task_hits"])
    np.testing.assert_array_equal(res.task_hitsum, exp["task_hitsum"])
    np.testing.assert_array_equal(res.task_status, exp["task_status"])
    if exact_scores:      # all statistics are exact integer sums + one IEEE division: bit-equal in practice
        np.testing.assert_array_equal(res.task_stat, exp["task_stat"])
        np.testing.assert_array_equal(res.task_score, exp["task_score"])
    else:
        np.testing.assert_allclose(res.task_score, exp["task_score"], rtol=0, atol=QS_TOL)
    np.testing.assert_array_equal(res.sv_nscore, exp["sv_nscore"])
    np.testing.assert_array_equal(res.sv_gt, exp["sv_gt"])
    np.testing.assert_allclose(res.sv_qs, exp["sv_qs"], rtol=0, atol=QS_TOL)
    np.testing.assert_allclose(res.sv_gs, exp["sv_gs"], rtol=0, atol=QS_TOL)
    np.testing.assert_allclose(res.sv_gq, exp["sv_gq"], rtol=0, atol=GQ_TOL)


def test_dotdata_bit_exact(engine):
    rng = np.random.default_rng(11)
    for k in (4, 10, 16, 20, 30, 40):
        ref = synth.random_dna(rng, 1500)
        reads, off = synth.simulate_reads(rng, ref, np.array([0]), np.array([1400]), np.array([1200]))
        got = engine.dotdata(k, reads, ref)
        exp = O.dotdata(k, reads.tobytes().decode(), ref.tobytes().decode())
        np.testing.assert_array_equal(got, exp)


def test_dotdata_palindromes_and_revcomp(engine):
    # ACGT is its own reverse complement: the reference emits those dots twice
    read, ref = "ACGTACGTAA", "ACGTTTACGT"
    got = engine.dotdata(4, read, ref)
    exp = O.dotdata(4, read, ref)
    np.testing.assert_array_equal(got, exp)
    assert (got[0] == got[1]).all()
    rng = np.random.default_rng(3)
    s = synth.random_dna(rng, 800)
    rc = synth.revcomp(s)
    for k in (10, 20):
        np.testing.assert_array_equal(engine.dotdata(k, rc, s), O.dotdata(k, rc.tobytes().decode(), s.tobytes().decode()))


def test_dotdata_alphabet_edge_cases(engine):
    rng = np.random.default_rng(4)
    s = synth.random_dna(rng, 600).copy()
    r = s.copy()
    s[100:140] = ord("N"); r[100:140] = ord("N")          # N k-mers match each other
    s[200:260] += 32                                        # lower-case structure never matches upper-case read
    s[300:320] = ord("X")                                   # X never matches
    s[400] = ord("R"); r[400] = ord("Y")                    # IUPAC -> N on both sides
    r[500:520] += 32; s[500:520] += 32                      # lower == lower
    for k in (10, 20):
        np.testing.assert_array_equal(engine.dotdata(k, r, s), O.dotdata(k, r.tobytes().decode(), s.tobytes().decode()))
    with pytest.raises(KeyError):
        bad = r.copy(); bad[50] = ord("X")
        engine.dotdata(10, bad, s)


def test_dotdata_short_and_empty(engine):
    assert len(engine.dotdata(10, "ACGT", "ACGTACGTACGTACGT")) == 0
    assert len(engine.dotdata(10, "ACGTACGTACGTACGT", "ACG")) == 0
    assert len(engine.dotdata(10, "", "")) == 0


def test_dotdata_repetitive_overflow(engine):
    # poly-A x poly-A: every cell matches -> far more hits than the first-pass capacity
    a = "A" * 700
    got = engine.dotdata(10, a, a)
    exp = O.dotdata(10, a, a)
    np.testing.assert_array_equal(got, exp)


def test_overflowing_waves_then_second_run(engine):
    """Repetitive plots (more dots than the first-pass capacity n + m + 32) in several waves of one resident batch,
    scored twice: the overflow re-run must not touch the resident plan (stale hit offsets after the
    overflow buffer was reallocated), so run() #2 equals run() #1 equals the oracle."""
    from vapor_b200.engine import Engine
    rng = np.random.default_rng(23)
    b = Batch()
    unit = synth.random_dna(rng, 37)
    for rep, extra in ((30, 0), (12, 1), (45, 2), (20, 0)):
        ref = np.concatenate([synth.random_dna(rng, 2500), np.tile(unit, rep), synth.random_dna(rng, 2500)])
        alt = np.concatenate([ref[:2600], np.tile(unit, rep + 8), ref[-2550:]])
        rid, aid = b.add_seq(ref), b.add_seq(alt)
        for j in range(3):
            read = np.concatenate([ref[: 2500 + 10 * j], np.tile(unit, rep + j), ref[-2500:]])
            b.add_task(b.add_seq(read), rid, aid, extra, 10, (MODE_ABS, MODE_W10, MODE_REDEF)[j])
        b.end_sv(rep)
    for _ in range(2):                                      # ordinary SVs between and after the repetitive ones
        case = synth.make_sv_case(rng, "DEL", 400, genotype=1)
        rid, aid = b.add_seq(case.ref_seq), b.add_seq(case.alt_seq)
        reads, _ = synth.simulate_reads(rng, case.hap_alt, np.array([0]), np.array([1100]), np.array([case.read_window]))
        b.add_task(b.add_seq(reads), rid, aid, 0, 10, MODE_ABS_AND_W10)
        b.end_sv("DEL")
    pb = b.pack()
    exp = BO.score_batch(pb)
    eng = Engine(0, hit_budget_bytes=1 << 16)               # smallest budget: about one task per wave
    try:
        eng.set_option("k2_mode", 1 if engine.k2_mode_name == "join" else 0)
        eng.upload(pb)
        eng.run()
        tm = eng.timings()
        assert tm["n_waves"] > 2 and tm["n_overflow_plots"] >= 4
        first = eng.fetch()
        eng.run()
        assert eng.timings()["n_overflow_plots"] == tm["n_overflow_plots"]
        second = eng.fetch()
    finally:
        eng.close()
    _compare(first, exp)
    for f in first.__dataclass_fields__:
        np.testing.assert_array_equal(getattr(second, f), getattr(first, f), err_msg=f)


def test_overflow_after_a_larger_batch(engine):
    """A plot that overflows its first-pass capacity is incomplete -- with self-reverse-complement k-mers (two slots per
    dot) the last slot before the capacity may never be written -- so kernel 3 must not look at it before the re-run:
    after a batch with larger coordinates the stale slab contents would index outside the small plot's bitmaps
    (found in round 2 as a timing-dependent illegal address)."""
    rng = np.random.default_rng(31)
    big = synth.make_workload(3, seed=5, types=("TANDUP",), size_range=(4000, 5000), reads_per_sv=4)
    engine.score(big.batch)                                   # fills the hit slab with coordinates up to ~11 000
    b = Batch()
    for rep in (40, 75, 120):                                 # ACGT repeats: every 10-mer window is one of four, two of them palindromic
        ref = np.concatenate([synth.random_dna(rng, 60), np.frombuffer(b"ACGT" * rep, np.uint8), synth.random_dna(rng, 60)])
        alt = np.concatenate([ref[:80], np.frombuffer(b"ACGT" * (rep // 2), np.uint8), ref[-70:]])
        rid, aid = b.add_seq(ref), b.add_seq(alt)
        for mode in (MODE_ABS, MODE_W10, MODE_REDEF, MODE_ABS_AND_W10):
            b.add_task(b.add_seq(ref[10:-10]), rid, aid, 0, 10, mode)
        b.end_sv(rep)
    pb = b.pack()
    exp = BO.score_batch(pb)
    for _ in range(4):
        res = engine.score(pb)
        assert engine.timings()["n_overflow_plots"] > 0
        _compare(res, exp)


def test_negative_miss_bp_slices_from_the_end(engine):
    """cigar2alignstart_by_pos can hand back a negative miss_bp; the reference then slices ref_seq[miss_bp:] the
    Python way (Simple_function.pyx:185-186).  The library does the same instead of rejecting the batch."""
    rng = np.random.default_rng(29)
    case = synth.make_sv_case(rng, "INV", 500, genotype=1)
    b = Batch()
    rid, aid = b.add_seq(case.ref_seq), b.add_seq(case.alt_seq)
    reads, _ = synth.simulate_reads(rng, case.hap_alt, np.array([0]), np.array([1300]), np.array([700]))
    for miss in (-700, -5, -100000, 0):
        b.add_task(b.add_seq(reads), rid, aid, miss, 10, MODE_ABS)
        b.add_task(b.add_seq(reads), rid, aid, miss, 10, MODE_W10)
    b.end_sv("neg")
    pb = b.pack()
    _compare(engine.score(pb), BO.score_batch(pb))


@pytest.mark.parametrize("seed,types", [(1, synth.SV_TYPES), (2, ("DEL",)), (3, ("TANDUP", "INV")), (4, ("INS",))])
def test_workload_parity(engine, seed, types):
    w = synth.make_workload(12, seed=seed, types=types, size_range=(50, 1500), reads_per_sv=8,
                            max_miss=3, lowercase_every=5, k_choices=(10, 10, 20, 10, 30))
    res = engine.score(w.batch)
    exp = BO.score_batch(w.batch)
    _compare(res, exp)
    assert (res.task_status == 1).sum() > 0


def test_modes_on_same_inputs(engine):
    """Every mode on every SV type (the drivers use each mode on several SV classes)."""
    rng = np.random.default_rng(9)
    b = Batch()
    for st in synth.SV_TYPES:
        case = synth.make_sv_case(rng, st, int(rng.integers(200, 1200)), genotype=1)
        ref_id, alt_id = b.add_seq(case.ref_seq), b.add_seq(case.alt_seq)
        for mode in (MODE_ABS, MODE_W10, MODE_REDEF, MODE_ABS_AND_W10):
            for hap in (case.hap_ref, case.hap_alt):
                want = case.read_window
                reads, _ = synth.simulate_reads(rng, hap, np.array([0]), np.array([min(len(hap), int(want * 1.12) + 60)]),
                                                np.array([want]))
                b.add_task(b.add_seq(reads), ref_id, alt_id, int(rng.integers(0, 4)), 10, mode)
            b.end_sv(st)
    pb = b.pack()
    res = engine.score(pb)
    _compare(res, BO.score_batch(pb))


def test_empty_and_unscorable(engine):
    b = Batch()
    rng = np.random.default_rng(5)
    ref = synth.random_dna(rng, 900); alt = synth.random_dna(rng, 700)
    rid, aid = b.add_seq(ref), b.add_seq(alt)
    b.add_task(b.add_seq(synth.random_dna(rng, 800)), rid, aid, 0, 10, MODE_ABS)      # unrelated read -> [0,0]
    b.add_task(b.add_seq("ACGT"), rid, aid, 0, 10, MODE_W10)                          # read shorter than k
    b.add_task(b.add_seq(""), rid, aid, 0, 10, MODE_REDEF)                            # empty read
    b.end_sv("none")
    b.end_sv("empty-sv")                                                              # SV without reads -> NA row
    pb = b.pack()
    res = engine.score(pb)
    _compare(res, BO.score_batch(pb))
    assert list(res.sv_gt) == [255, 255]


def test_hit_budget_waves_identical(engine):
    from vapor_b200.engine import Engine
    w = synth.make_workload(10, seed=21, size_range=(50, 800), reads_per_sv=6)
    base = engine.score(w.batch)
    small = Engine(0, hit_budget_bytes=1 << 20)         # force many waves
    try:
        small.set_option("k2_mode", 1 if engine.k2_mode_name == "join" else 0)
        res = small.score(w.batch)
        assert small.timings()["n_waves"] > 1
    finally:
        small.close()
    for f in base.__dataclass_fields__:
        np.testing.assert_array_equal(getattr(res, f), getattr(base, f))


def test_selfplot_qc_counts(engine):
    """vapor_gpu_selfplot_qc against the counts qual_check_repetitive_region takes from dotdata(k, s, s)
    (Simple_function.pyx:1154-1171), on random, repetitive, palindromic and N-holding windows."""
    rng = np.random.default_rng(17)
    seqs, ks = [], []
    base = synth.random_dna(rng, 2600)
    seqs.append(base); ks.append(10)
    rep = np.concatenate([base[:700], base[200:700], base[200:700], base[700:1500]])      # tandem repeats
    seqs.append(rep); ks.append(10)
    seqs.append(rep); ks.append(20)
    seqs.append(rep); ks.append(40)
    inv = np.concatenate([base[:900], synth.revcomp(base[300:900]), base[900:1200]])       # inverted repeat
    seqs.append(inv); ks.append(10)
    withn = base[:1500].copy(); withn[300:350] = ord("N")
    seqs.append(withn); ks.append(10)
    seqs.append(np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGT", dtype=np.uint8)); ks.append(4)   # palindromic k-mers
    seqs.append(np.frombuffer(b"ACGTA", dtype=np.uint8)); ks.append(10)                    # shorter than k
    seqs.append(synth.random_dna(rng, 6000)); ks.append(10)                                # several strips, transposed tail
    got = engine.selfplot_qc(seqs, ks)
    for i, (s, k) in enumerate(zip(seqs, ks)):
        st = s.tobytes().decode()
        d = O.dotdata(k, st, st)
        low = d[d[:, 0] > d[:, 1]]
        exp = [len(d), int((d[:, 0] == d[:, 1]).sum()), len(low)]
        assert got[i, :3].tolist() == exp, (i, k)
        if len(low):
            assert got[i, 3:7].tolist() == [low[:, 0].min(), low[:, 0].max(), low[:, 1].min(), low[:, 1].max()], i
        assert got[i, 7] == 1
    bad = base[:400].copy(); bad[77] = ord("X")
    assert engine.selfplot_qc([bad], [10])[0, 7] == 2


@pytest.mark.parametrize("svtype,svlen,k,mode", [
    ("INV", 12000, 10, MODE_ABS), ("TANDUP", 30000, 20, MODE_REDEF), ("DEL", 60000, 30, MODE_W10),
    ("INV", 100000, 40, MODE_ABS), ("INS", 20000, 10, MODE_W10), ("TANDUP", 11000, 10, MODE_ABS_AND_W10)])
def test_large_windows(engine, svtype, svlen, k, mode):
    """BASELINE config 4: 10-100 kb windows.  The reference drivers never plot that much (they fall back to 1 kb
    junction windows at >= 10 kb, Simple_function.pyx:1728), but its scoring functions accept any length; these
    plots span many strips on both axes, transposed tails, several kernel-3 scratch classes and every k."""
    rng = np.random.default_rng(svlen + k)
    case = synth.make_sv_case(rng, svtype, svlen, genotype=1)
    b = Batch()
    rid, aid = b.add_seq(case.ref_seq), b.add_seq(case.alt_seq)
    for hap, miss in ((case.hap_alt, 0), (case.hap_ref, 3)):
        want = case.read_window - miss
        reads, _ = synth.simulate_reads(rng, hap, np.array([miss]), np.array([min(len(hap) - miss, int(want * 1.12) + 60)]),
                                        np.array([want]))
        b.add_task(b.add_seq(reads), rid, aid, miss, k, mode)
    b.end_sv(svtype)
    pb = b.pack()
    res = engine.score(pb)
    _compare(res, BO.score_batch(pb))


def test_properties_at_benchmark_scale(engine):
    """Size-independent properties on a batch too large for the oracle to score in full (BASELINE config 2 sizes):
    (1) waves: a 64 MB hit budget (many waves) gives byte-identical results to one wave; (2) sharding: scoring the
    SVs in two halves and merging equals scoring them together; (3) the kernel-2 variants (join kernel, all-pairs tile
    kernel with the ISETP-only and the dual-pipe inner loop) give identical hit counts and coordinate checksums;
    (4) a sample of SVs agrees with the oracle."""
    from vapor_b200 import multi
    from vapor_b200.engine import Engine
    w = synth.make_workload(120, seed=77, size_range=(50, 5000), reads_per_sv=20)
    base = engine.score(w.batch)
    assert (base.task_status == 1).sum() > 0.5 * w.batch.n_task
    small = Engine(0, hit_budget_bytes=64 << 20)
    try:
        small.set_option("k2_mode", 1 if engine.k2_mode_name == "join" else 0)
        waves = small.score(w.batch)
        assert small.timings()["n_waves"] > 1
        small.set_option("k2_mode", 0)
        small.set_option("tile_variant", 0)
        v0 = small.score(w.batch)
        small.set_option("tile_variant", 4)
        v4 = small.score(w.batch)
        small.set_option("k2_mode", 1)
        vj = small.score(w.batch)
        tmj = small.timings()
        assert tmj["k2_mode"] == 1 and 0 < tmj["evaluated_cells"] < tmj["cells"] // 100
    finally:
        small.close()
    for f in base.__dataclass_fields__:
        for other in (waves, v0, v4, vj):
            np.testing.assert_array_equal(getattr(other, f), getattr(base, f), err_msg=f)
    import threading
    lock = threading.Lock()                                  # one handle is not re-entrant: serialise the two shards

    def one_at_a_time(pb):
        with lock:
            return engine.score(pb)
    merged = multi.score_sharded(w.batch, [one_at_a_time, one_at_a_time])
    for f in base.__dataclass_fields__:
        np.testing.assert_array_equal(getattr(merged, f), getattr(base, f), err_msg=f)
    sample = w.batch.shard([3, 17, 58, 111])
    exp = BO.score_batch(sample)
    got = engine.score(sample)
    _compare(got, exp)


def test_pipeline_double_buffered(engine):
    """engine.Pipeline: several batches through two handles give the same results as one handle, in order."""
    from vapor_b200.engine import Pipeline
    ws = [synth.make_workload(6, seed=100 + i, size_range=(50, 900), reads_per_sv=5) for i in range(5)]
    exp = [engine.score(w.batch) for w in ws]
    with Pipeline(0, depth=2) as pipe:
        got = pipe.map([w.batch for w in ws])
    for g, e in zip(got, exp):
        for f in e.__dataclass_fields__:
            np.testing.assert_array_equal(getattr(g, f), getattr(e, f), err_msg=f)


def test_kernel3_variants_and_overlap_agree(engine):
    """Kernel 3 has two implementations (one warp per task, k3_warp.cuh; one CTA per task, k3_score.cuh) and the run can
    put kernel 3 of a wave on a second stream under kernels 1-2 of the next wave: every combination gives byte-identical
    results on a workload of every simple SV type, mode and k, lower case and miss_bp > 0 included -- with all five
    scratch classes on the warp kernel, with none, and with the default split -- and a sample agrees with the oracle."""
    from vapor_b200.engine import Engine
    w = synth.make_workload(160, seed=4711, size_range=(50, 5000), reads_per_sv=12, max_miss=4, lowercase_every=7,
                            k_choices=(10, 10, 20, 10, 30, 10, 40))
    base = engine.score(w.batch)
    other = Engine(0, hit_budget_bytes=48 << 20)            # several waves, so that the overlap has something to overlap
    try:
        other.set_option("k2_mode", 1 if engine.k2_mode_name == "join" else 0)
        for opts in ({"k3_mode": 0}, {"k3_mode": 1, "k3_warp_classes": 5}, {"k3_mode": 1, "k3_warp_classes": 1},
                     {"k3_mode": 1, "k3_warp_classes": 3, "overlap": 1}, {"k3_mode": 0, "overlap": 1}):
            for k_, v_ in opts.items():
                other.set_option(k_, v_)
            got = other.score(w.batch)
            assert other.timings()["n_waves"] > 1
            for f in base.__dataclass_fields__:
                np.testing.assert_array_equal(getattr(got, f), getattr(base, f), err_msg=f"{opts} {f}")
            again = other.score(w.batch)                     # the same handle, the same resident buffers, a second time
            for f in base.__dataclass_fields__:
                np.testing.assert_array_equal(getattr(again, f), getattr(base, f), err_msg=f"{opts} second run {f}")
    finally:
        other.close()
    sample = w.batch.shard([1, 2, 3, 4, 50, 99, 140])
    _compare(engine.score(sample), BO.score_batch(sample))


def test_short_range_plot_with_more_than_65535_dots(engine):
    """A tandem repeat inside a SHORT window: the value range (n + m - 1 = 5 000 bins) puts the task in a class of the
    warp-per-task kernel 3, whose group sizes are 16-bit counters, but the plot holds ~160 000 dots.  Such a plot overflows
    its first-pass slab, and the re-scoring of its task must go to the CTA-per-task kernel (32-bit counters): results equal
    the oracle in every mode."""
    rng = np.random.default_rng(99)
    b = Batch()
    unit = synth.random_dna(rng, 11)
    left, right = synth.random_dna(rng, 600), synth.random_dna(rng, 600)
    ref = np.concatenate([left, np.tile(unit, 120), right])
    alt = np.concatenate([left, np.tile(unit, 100), right])
    rid, aid = b.add_seq(ref), b.add_seq(alt)
    for j, mode in enumerate((MODE_ABS, MODE_W10, MODE_REDEF, MODE_ABS_AND_W10)):
        read = np.concatenate([left[5 * j:], np.tile(unit, 118 + j), right])
        b.add_task(b.add_seq(read), rid, aid, 0, 10, mode)
    b.end_sv("repeat")
    pb = b.pack()
    exp = BO.score_batch(pb)
    assert int(exp["task_hits"].max()) > 65535
    got = engine.score(pb)
    assert engine.timings()["n_overflow_plots"] > 0
    _compare(got, exp)
