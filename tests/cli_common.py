"""Shared checks of the end-to-end CLI tests: run ``vapor bed`` / ``vapor vcf`` on the committed synthetic case and
compare with the tables the unmodified reference CLI wrote (tests/golden/make_cli_golden.py)."""
import json
import os
import shutil

from vapor_b200 import Simple_function as SF
from vapor_b200 import cli

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = os.path.join(HERE, "golden", "cli_case")
QS_TOL = 1e-5     # north_star: VaPoR_qs/gs/Rec within 1e-5 absolute
GQ_TOL = 1e-3     # north_star: VaPoR_GQ within 1e-3


class Args:
    def __init__(self, **kw):
        self.PB_supp = None
        self.gpus = 1
        self.__dict__.update(kw)


def _rows(path):
    with open(path) as f:
        return [line.rstrip("\n").split("\t") for line in f]


def compare_bed_tables(got_path, exp_path):
    got, exp = _rows(got_path), _rows(exp_path)
    assert got[0] == exp[0]
    assert len(got) == len(exp)
    for g, e in zip(got[1:], exp[1:]):
        assert g[:5] == e[:5], (g, e)
        if e[5] == "NA":
            assert g[5:] == e[5:]
            continue
        assert abs(float(g[5]) - float(e[5])) <= QS_TOL, (g, e)      # QS
        assert abs(float(g[6]) - float(e[6])) <= QS_TOL, (g, e)      # GS
        assert g[7] == e[7], (g, e)                                   # GT identical
        assert abs(float(g[8]) - float(e[8])) <= GQ_TOL, (g, e)      # GQ
        assert g[9] == e[9], (g, e)                                   # Rec: rounded strings, identical


def _info_fields(info):
    out = {}
    for kv in info.split(";"):
        if "=" in kv:
            k, v = kv.split("=", 1)
            out[k] = v
    return out


def compare_annotated_vcf(got_path, exp_path):
    got, exp = _rows(got_path), _rows(exp_path)
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        if len(e) < 8:
            assert g == e
            continue
        assert g[:7] == e[:7] and g[8:] == e[8:], (g, e)
        gi, ei = _info_fields(g[7]), _info_fields(e[7])
        assert gi.keys() == ei.keys()
        for k in ei:
            if k == "VaPor_GQ" and ei[k] != "NA":
                assert abs(float(gi[k]) - float(ei[k])) <= 0.011, (k, gi[k], ei[k])    # rounded to 2 dp in the file
            else:
                assert gi[k] == ei[k], (k, gi[k], ei[k])


CASE_LARGE = os.path.join(HERE, "golden", "cli_case_large")


def run_bed_case(tmp_path, session, case=None):
    case = case or CASE
    out = os.path.join(str(tmp_path), "bed.vapor")
    args = Args(sv_input=os.path.join(case, "svs.bed"), output_path=os.path.join(str(tmp_path), "figs"), output_file=out,
                reference=os.path.join(case, "ref.fa"), pacbio_input=os.path.join(case, "reads.sam.gz"))
    SF.set_session(session)
    try:
        cli.run_bed(args, [session])
    finally:
        SF.set_session(None)
    compare_bed_tables(out, os.path.join(case, "svs.bed.vapor.golden"))


CASE_COMPLEX = os.path.join(HERE, "golden", "cli_case_complex")


def run_vcf_case(tmp_path, session, case=None, first_records=0):
    """``first_records`` > 0: only the first records of the VCF (and of the golden table, which keeps the VCF's order)."""
    case = case or CASE
    vcf = os.path.join(str(tmp_path), "svs_nohdr.vcf")
    gold = os.path.join(case, "svs_nohdr.vcf.vapor.golden")
    if first_records:
        with open(os.path.join(case, "svs_nohdr.vcf")) as f, open(vcf, "w") as g:
            g.writelines(f.readlines()[:first_records])
        gold_cut = os.path.join(str(tmp_path), "golden_cut.vapor")
        with open(gold) as f, open(gold_cut, "w") as g:
            lines = f.readlines()
            head = 1 if lines and not lines[0].strip() else 0         # the reference writes an empty header block first
            g.writelines(lines[:first_records + head])
        gold = gold_cut
    else:
        shutil.copy(os.path.join(case, "svs_nohdr.vcf"), vcf)
    args = Args(sv_input=vcf, output_path=os.path.join(str(tmp_path), "figs"), output_file="unused",
                reference=os.path.join(case, "ref.fa"), pacbio_input=os.path.join(case, "reads.sam.gz"))
    SF.set_session(session)
    try:
        cli.run_vcf(args, [session])
    finally:
        SF.set_session(None)
    compare_annotated_vcf(vcf + ".vapor", gold)


def run_disdup_case(session):
    gold = json.load(open(os.path.join(CASE, "disdup.golden.json")))
    SF.set_session(session)
    try:
        for svid, rec in gold.items():
            got = SF.vapor_simple_disdup_Vapor(3, 1, os.path.join(CASE, "reads.sam.gz"), os.path.join(CASE, "ref.fa"),
                                               list(rec["sv_info"]), "/tmp/unused.png")
            assert len(got) == len(rec["scores"]), svid
            for a, b in zip(got, rec["scores"]):
                assert abs(a - b) <= QS_TOL, (svid, a, b)
    finally:
        SF.set_session(None)
