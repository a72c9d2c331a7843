"""The CPU oracle against the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py).  Runs anywhere (no GPU, no /root/reference)."""
import json
import os

import numpy as np
import pytest

from oracle import batch_oracle as BO
from oracle import vapor_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "scoring_cases.json")))["cases"]
SUMM = json.load(open(os.path.join(HERE, "golden", "sv_summaries.json")))["summaries"]
MODES = ["calcu_vapor_single_read_score_abs_dis_m1b", "calcu_vapor_single_read_score_within_10Perc_m1b",
         "calcu_vapor_single_read_score_directed_dis_m1b_redefine_diagnal"]


def _check_dots(d, rec):
    assert len(d) == rec["n"]
    assert str(BO.hit_checksum(d)) == rec["checksum"]
    assert d[:40].tolist() == rec["head"]
    assert d[-10:].tolist() == rec["tail"] or rec["n"] == 0


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_scoring_case(case):
    read, ref, alt, miss, k = case["read"], case["ref"], case["alt"], case["miss"], case["k"]
    if case.get("error") == "KeyError":
        with pytest.raises(KeyError):
            O.dotdata(k, read, ref[miss:])
        return
    _check_dots(O.dotdata(k, read, ref[miss:]), case["dot_ref"])
    _check_dots(O.dotdata(k, read, alt[miss:]), case["dot_alt"])
    _check_dots(O.dotdata(k, read, ref.upper()[miss:]), case["dot_ref_upper"])
    for m in MODES:
        got = getattr(O, m)(ref, alt, [read, miss, "q"], k)
        assert [float(got[0]), float(got[1])] == case[m], m           # bit-equal
    d = O.dotdata(k, read, ref[miss:])
    if len(d):
        c = O.clean_dotdata_diagnal_and_anti_diagnal(d)
        assert len(c) == case["n_clean_a6"]
        assert len(O.clean_dotdata_diagnal_m1b(d)) == case["n_clean_diag_m1b"]
        assert len(O.clean_dotdata_anti_diagnal_m1b(d)) == case["n_clean_anti_m1b"]
        if len(c):
            assert float(O.dis_to_diagnal_most_abundant_defined(c)) == case["intercept"]


def test_sv_summaries():
    for s in SUMM:
        if not s["scores"]:
            assert O.result_organize_ins(["key", []])[1:] == s["row"]
            assert O.summarize_sv([]) is None
            continue
        row = O.result_organize_ins(["key", s["scores"]])
        assert float(row[1]) == s["QS"] and float(row[2]) == s["GS"] and row[3] == s["Rec"]
        gt, gq = O.gt_estimate_log_likelihood(row)
        assert gt == s["GT"]
        assert abs(float(gq) - s["GQ"]) <= 1e-12
