"""End-to-end on the GPU: ``vapor bed`` / ``vapor vcf`` through the real CUDA engine must reproduce the tables the
unmodified reference CLI wrote for the committed synthetic case (tests/golden/make_cli_golden.py), and the drop-in
single-call functions must agree with the oracle."""
import os

import numpy as np
import pytest

from oracle import vapor_oracle as O
from vapor_b200 import Simple_function as SF
from vapor_b200 import synth

import cli_common as CC

pytestmark = pytest.mark.gpu


@pytest.fixture()
def session(engine):
    s = SF.Session(engine=engine)
    SF.set_session(s)
    yield s
    SF.set_session(None)


def test_bed_cli_matches_reference_golden(tmp_path, session):
    CC.run_bed_case(tmp_path, session)
    assert session.stats["reads_scored"] > 50


def test_bed_cli_large_events_match_reference_golden(tmp_path, session):
    """Events >= 10 kb: the drivers' junction-window fallbacks (W10 mode) and the long-insertion window, against the
    table the unmodified reference CLI wrote for tests/golden/cli_case_large."""
    CC.run_bed_case(tmp_path, session, CC.CASE_LARGE)


def test_vcf_cli_matches_reference_golden(tmp_path, session):
    CC.run_vcf_case(tmp_path, session)


def test_vcf_cli_complex_events_match_reference_golden(tmp_path, session):
    """All 56 complex events of tests/golden/cli_case_complex (DEL_INV, DUP_INV, DEL_DUP_INV, swapped blocks, two- and
    three-allele `Other=` records, junction-window fallbacks of >= 10 kb events) against the reference CLI's annotated VCF."""
    CC.run_vcf_case(tmp_path, session, CC.CASE_COMPLEX)


def test_disdup_driver_matches_reference_golden(session):
    CC.run_disdup_case(session)


def test_dropin_single_calls(session):
    rng = np.random.default_rng(23)
    case = synth.make_sv_case(rng, "INV", 700, genotype=1)
    ref, alt = case.ref_seq.tobytes().decode(), case.alt_seq.tobytes().decode()
    reads, _ = synth.simulate_reads(rng, case.hap_alt, np.array([0]), np.array([int(case.read_window * 1.12) + 60]),
                                    np.array([case.read_window]))
    x = [reads.tobytes().decode(), 2, "q1"]
    assert SF.dotdata(10, x[0], ref) == [tuple(r) for r in O.dotdata(10, x[0], ref).tolist()]
    for name in ("calcu_vapor_single_read_score_abs_dis_m1b", "calcu_vapor_single_read_score_within_10Perc_m1b",
                 "calcu_vapor_single_read_score_directed_dis_m1b_redefine_diagnal"):
        got = getattr(SF, name)(ref, alt, x, 10)
        exp = getattr(O, name)(ref, alt, x, 10)
        assert [float(v) for v in got] == [float(v) for v in exp], name
    assert SF.window_size_refine(ref)[0] == 10
    assert SF.window_size_refine("ACGT")[0] == "Error"
    assert SF.window_size_refine("N" * 150 + ref)[0] == "Error"
    row = SF.result_organize_ins(["k", [0.5, -0.2, 0.004, 0.9]])
    exp = O.result_organize_ins(["k", [0.5, -0.2, 0.004, 0.9]])
    assert row[0] == exp[0] and abs(row[1] - exp[1]) < 1e-12 and row[2] == exp[2] and row[3] == exp[3]
    gt, gq = SF.gt_estimate_log_likelihood(row)
    egt, egq = O.gt_estimate_log_likelihood(exp)
    assert gt == egt and abs(gq - egq) <= 1e-3
    assert SF.result_organize_ins(["k", []]) == ["k", "NA", "NA", "NA"]


def test_two_sessions_same_output(tmp_path, engine):
    """Sharding the events over two sessions (here two handles on the same GPU) gives byte-identical tables."""
    from vapor_b200 import cli
    from vapor_b200.engine import Engine
    outs = []
    for n in (1, 2):
        sessions = [SF.Session(engine=engine)] + [SF.Session(engine=Engine(0)) for _ in range(n - 1)]
        out = os.path.join(str(tmp_path), f"bed_{n}.vapor")
        args = CC.Args(sv_input=os.path.join(CC.CASE, "svs.bed"), output_path=os.path.join(str(tmp_path), "figs"), output_file=out,
                       reference=os.path.join(CC.CASE, "ref.fa"), pacbio_input=os.path.join(CC.CASE, "reads.sam.gz"))
        SF.set_session(sessions[0])
        try:
            cli.run_bed(args, sessions)
        finally:
            SF.set_session(None)
            for s in sessions[1:]:
                s.close()
        outs.append(open(out).read())
    assert outs[0] == outs[1]


def test_worker_processes_same_output(tmp_path):
    """`--gpus N` runs one worker process per GPU (here 2 workers sharing the one GPU): byte-identical table."""
    import subprocess
    import sys
    root = os.path.dirname(CC.HERE)
    outs = []
    for n in (1, 2):
        out = os.path.join(str(tmp_path), f"bed_p{n}.vapor")
        subprocess.run([sys.executable, os.path.join(root, "bin", "vapor"), "bed", "--sv-input", os.path.join(CC.CASE, "svs.bed"),
                        "--output-path", os.path.join(str(tmp_path), "figs"), "--output-file", out,
                        "--reference", os.path.join(CC.CASE, "ref.fa"), "--pacbio-input", os.path.join(CC.CASE, "reads.sam.gz"),
                        "--gpus", str(n)], check=True, stdout=subprocess.DEVNULL)
        outs.append(open(out).read())
    assert outs[0] == outs[1]
    CC.compare_bed_tables(os.path.join(str(tmp_path), "bed_p2.vapor"), os.path.join(CC.CASE, "svs.bed.vapor.golden"))


def test_figure_hook_writes_dot_lists(tmp_path, session, monkeypatch):
    """VAPOR_FIGURES=tsv: the four recurrence plots the reference draws per event (ref/ref, alt/alt, best read/ref,
    best read/alt; Simple_function.pyx:1072-1089) are written as dot lists computed on the GPU."""
    monkeypatch.setenv("VAPOR_FIGURES", "tsv")
    fig = os.path.join(str(tmp_path), "ev.png")
    scores = SF.vapor_simple_inv_Vapor(3, 1, os.path.join(CC.CASE, "reads.sam.gz"), os.path.join(CC.CASE, "ref.fa"),
                                       ["chr1", 24714, 25480], fig)
    assert len(scores) > 3
    rows = [l.split("\t") for l in open(fig + ".dots.tsv")]
    names = {r[0] for r in rows}
    assert names == {"ref_vs_ref", "alt_vs_alt", "read_vs_ref", "read_vs_alt"}
    diag = [r for r in rows if r[0] == "ref_vs_ref" and r[1] == r[2].strip()]
    assert len(diag) > 1000                                   # the self-plot holds the whole diagonal


def test_integration_md_ctypes_stub_runs(engine):
    """The ctypes stub INTEGRATION.md shows a maintainer of the reference is executed as written (only the library
    path is made absolute) and must give the same score list and summary as the engine."""
    import re
    from vapor_b200 import _native
    from vapor_b200.engine import Batch, MODE_ABS
    root = os.path.dirname(CC.HERE)
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# vapor_vali/_b200\.py.*?)```", md, re.S).group(1)
    block = block.replace('C.CDLL("libvapor_b200.so")', f'C.CDLL({_native.LIB_PATH!r})')
    ns = {}
    exec(compile(block, "INTEGRATION.md:_b200.py", "exec"), ns)
    rng = np.random.default_rng(31)
    case = synth.make_sv_case(rng, "INV", 600, genotype=1)
    ref, alt = case.ref_seq.tobytes().decode(), case.alt_seq.tobytes().decode()
    reads = []
    for hap in (case.hap_alt, case.hap_ref, case.hap_alt, case.hap_alt):
        r, _ = synth.simulate_reads(rng, hap, np.array([0]), np.array([int(case.read_window * 1.12) + 60]), np.array([case.read_window]))
        reads.append([r.tobytes().decode(), 0, "q"])
    scores, qs, gs, gt, gq = ns["score_reads"](ref, alt, reads, 10, ns["MODE_ABS"])
    b = Batch()
    rid, aid = b.add_seq(ref), b.add_seq(alt)
    for x in reads:
        b.add_task(b.add_seq(x[0]), rid, aid, 0, 10, MODE_ABS)
    b.end_sv("e")
    pb = b.pack()
    res = engine.score(pb)
    assert scores == res.sv_scores(pb, 0) and len(scores) >= 3
    assert qs == res.sv_qs[0] and gs == res.sv_gs[0] and gt == int(res.sv_gt[0]) and gq == res.sv_gq[0]


def test_figure_hook_writes_png(tmp_path, session, monkeypatch):
    """VAPOR_FIGURES=png: a 2 x 2 recurrence-plot image per event (matplotlib if present, else the built-in writer)."""
    monkeypatch.setenv("VAPOR_FIGURES", "png")
    fig = os.path.join(str(tmp_path), "ev.png")
    SF.vapor_simple_del_Vapor(3, 1, os.path.join(CC.CASE, "reads.sam.gz"), os.path.join(CC.CASE, "ref.fa"), ["chr1", 12000, 12643], fig)
    data = open(fig, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and len(data) > 1000


def test_kselect_on_repeat_rich_windows_matches_reference_golden(session):
    """The same 48 repeat-rich windows as the CPU suite, the self-plots and the below-diagonal dot lists from the GPU."""
    import json
    import os
    import numpy as np
    cases = json.load(open(os.path.join(CC.HERE, "golden", "kselect_cases.json")))["cases"]
    for c in cases:
        np.random.seed(c["seed"])
        got = session.refine_many([SF.RefineRequest(c["seq"])])[0][0]
        assert got == c["k"], (c["kind"], len(c["seq"]), got, c["k"])
