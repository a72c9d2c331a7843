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
