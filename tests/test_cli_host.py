"""Host logic of the drop-in layer, end to end, WITHOUT a GPU: the command line runs on the committed synthetic
case with the oracle standing in for the CUDA engine (tests/oracle_engine.py, test infrastructure) and must
reproduce the tables the unmodified reference CLI wrote.  This pins window arithmetic, read chopping, the driver
control flow, the coroutine scheduler and the writers; the GPU suite repeats it with the real engine."""
import os

import pytest

from vapor_b200 import Simple_function as SF
from vapor_b200 import cli

import cli_common as CC
from oracle_engine import OracleEngine


@pytest.fixture()
def session():
    s = SF.Session(engine=OracleEngine())
    yield s


def test_bed_cli_matches_reference_golden(tmp_path, session):
    CC.run_bed_case(tmp_path, session)
    assert session.stats["rounds"] <= 6                    # a handful of batched rounds, not one per SV


def test_bed_cli_large_events_match_reference_golden(tmp_path, session):
    """Events >= 10 kb: the drivers' junction-window fallbacks (W10 mode) and the long-insertion window, against the
    table the unmodified reference CLI wrote for tests/golden/cli_case_large."""
    CC.run_bed_case(tmp_path, session, CC.CASE_LARGE)


def test_vcf_cli_matches_reference_golden(tmp_path, session):
    CC.run_vcf_case(tmp_path, session)


def test_vcf_cli_complex_events_match_reference_golden(tmp_path, session):
    (tmp_path / "cut").mkdir()
    """tests/golden/cli_case_complex: DEL_INV, DUP_INV, DEL_DUP_INV (two alternative haplotypes), swapped blocks, two- and
    three-allele `Other=` records and a >= 10 kb event on the junction-window fallback, against the annotated VCF the
    unmodified reference CLI wrote (56 records; the GPU suite repeats it with the real engine)."""
    CC.run_vcf_case(tmp_path, session, CC.CASE_COMPLEX)
    CC.run_vcf_case(tmp_path / "cut", session, CC.CASE_COMPLEX, first_records=3)      # a prefix of the file gives a prefix of the table


def test_disdup_driver_matches_reference_golden(session):
    CC.run_disdup_case(session)


def test_vcf_with_header_annotates_the_right_records(tmp_path, session):
    """The reference mis-numbers records when the VCF has header lines; here the annotation must land on the same
    records, with the same values, as for the header-less copy."""
    import shutil
    vcf = os.path.join(str(tmp_path), "svs.vcf")
    with open(os.path.join(CC.CASE, "svs.vcf")) as f, open(vcf, "w") as g:
        for line in f:
            if "<DISDUP>" not in line:
                g.write(line)
    args = CC.Args(sv_input=vcf, output_path=os.path.join(str(tmp_path), "figs"), output_file="unused",
                   reference=os.path.join(CC.CASE, "ref.fa"), pacbio_input=os.path.join(CC.CASE, "reads.sam.gz"))
    SF.set_session(session)
    try:
        cli.run_vcf(args, [session])
    finally:
        SF.set_session(None)
    got = [l.rstrip("\n") for l in open(vcf + ".vapor")]
    body = [l for l in got if not l.startswith("#")]
    exp = [l.rstrip("\n") for l in open(os.path.join(CC.CASE, "svs_nohdr.vcf.vapor.golden")) if l.strip()]
    assert len(body) == len(exp)
    for g, e in zip(body, exp):
        assert g.split("\t")[:7] == e.split("\t")[:7]
        assert ("VaPor_GT=" in g) == ("VaPor_GT=" in e)
    assert got[0].startswith("##fileformat")
    assert sum(1 for l in got if l.startswith("##INFO=<ID=VaPoR_")) == 4


def test_svelter_subcommand_matches_vcf_other_path(tmp_path, session):
    """`vapor svelter` (vapor_vali/vapor:467-492) drives the same CANNOT_CLASSIFY driver as the VCF 'Other=' records:
    the same event through both entry points gives the same scores."""
    import json
    truth = json.load(open(os.path.join(CC.CASE, "truth.json")))
    ev = [t for t in truth if t["type"] == "OTHER"][0]
    b0, b1, b2 = ev["bps"]
    sv = os.path.join(str(tmp_path), "calls.svelter")
    with open(sv, "w") as f:
        f.write("chr\tstart\tend\tbp_info\tref\talt\n")
        f.write(f"chr1\t{b0}\t{b2}\tchr1:{b0}:{b1}:{b2}\tab/ab\tab/ba\n")
    out = os.path.join(str(tmp_path), "svelter.vapor")
    args = CC.Args(sv_input=sv, output_path=os.path.join(str(tmp_path), "figs"), output_file=out,
                   reference=os.path.join(CC.CASE, "ref.fa"), pacbio_input=os.path.join(CC.CASE, "reads.sam.gz"))
    SF.set_session(session)
    try:
        cli.run_svelter(args, [session])
    finally:
        SF.set_session(None)
    row = open(out).read().strip().split("\t")
    gold = [l.rstrip("\n").split("\t") for l in open(os.path.join(CC.CASE, "svs_nohdr.vcf.vapor.golden")) if "<OTHER>" in l][0]
    rec = [kv for kv in gold[7].split(";") if kv.startswith("VaPor_REC=")][0].split("=", 1)[1]
    assert row[0] == f".chr1_{b0}_{b1}_{b2}" and row[-1] == rec


def test_kselect_on_repeat_rich_windows_matches_reference_golden(session):
    """window_size_refine on windows with planted tandem / inverted / interspersed repeats: the k the unmodified reference
    chose (tests/golden/kselect_cases.json, numpy's random state seeded as recorded -- its X-means is otherwise unseeded;
    36 of the 48 windows reach the X-means sizing of the below-diagonal dots, Simple_function.pyx:1154-1171, 856-906,
    2101-2116).  profiles/r02_kselect_agreement.json holds the same comparison on 2 000 windows (100 % agreement)."""
    import json
    import numpy as np
    cases = json.load(open(os.path.join(CC.HERE, "golden", "kselect_cases.json")))["cases"]
    assert sum(1 for c in cases if c["xmeans_branch_at_k10"]) >= 30
    for c in cases:
        np.random.seed(c["seed"])
        got = session.refine_many([SF.RefineRequest(c["seq"])])[0][0]
        assert got == c["k"], (c["kind"], len(c["seq"]), got, c["k"])
