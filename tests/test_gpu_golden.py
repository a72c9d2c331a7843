"""The CUDA path against the committed golden vectors generated from the unmodified reference."""
import json
import os

import numpy as np
import pytest

from vapor_b200.engine import Batch, MODE_ABS, MODE_REDEF, MODE_W10, hit_checksum

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "scoring_cases.json")))["cases"]
SUMM = json.load(open(os.path.join(HERE, "golden", "sv_summaries.json")))["summaries"]
MODES = [(MODE_ABS, "calcu_vapor_single_read_score_abs_dis_m1b"), (MODE_W10, "calcu_vapor_single_read_score_within_10Perc_m1b"),
         (MODE_REDEF, "calcu_vapor_single_read_score_directed_dis_m1b_redefine_diagnal")]


def test_dotdata_golden(engine):
    for c in CASES:
        if c.get("error"):
            with pytest.raises(KeyError):
                engine.dotdata(c["k"], c["read"], c["ref"][c["miss"]:])
            continue
        for key, struct in (("dot_ref", c["ref"][c["miss"]:]), ("dot_alt", c["alt"][c["miss"]:]),
                            ("dot_ref_upper", c["ref"].upper()[c["miss"]:])):
            d = engine.dotdata(c["k"], c["read"], struct)
            rec = c[key]
            assert len(d) == rec["n"], (c["name"], key)
            assert str(hit_checksum(d)) == rec["checksum"], (c["name"], key)
            assert d[:40].tolist() == rec["head"], (c["name"], key)


def test_modes_golden(engine):
    b = Batch()
    idx = []
    for c in CASES:
        if c.get("error"):
            continue
        r, f, a = b.add_seq(c["read"]), b.add_seq(c["ref"]), b.add_seq(c["alt"])
        for mode, name in MODES:
            idx.append((b.add_task(r, f, a, c["miss"], c["k"], mode), c, name))
        b.end_sv(c["name"])
    res = engine.score(b.pack())
    for t, c, name in idx:
        assert res.task_stat[t, :2].tolist() == c[name], (c["name"], name)     # bit-equal pairs
        ref_hits = c["dot_ref_upper"]["n"] if name.endswith("abs_dis_m1b") else c["dot_ref"]["n"]
        assert int(res.task_hits[t, 0]) == ref_hits, (c["name"], name)


def test_bad_read_status(engine):
    c = [c for c in CASES if c.get("error")][0]
    b = Batch()
    b.add_task(b.add_seq(c["read"]), b.add_seq(c["ref"]), b.add_seq(c["alt"]), 0, c["k"], MODE_ABS)
    b.end_sv("bad")
    res = engine.score(b.pack())
    assert res.task_status[0] == 2 and res.sv_gt[0] == 255


def test_sv_summaries_golden(engine):
    """Kernel 4 alone (vapor_gpu_summarize) against result_organize_ins / gt_estimate_log_likelihood of the
    reference on 120 score vectors, including scores around the round(x, 2) edge at 0.005."""
    lists = [s["scores"] for s in SUMM]
    qs, gs, gq, gt, ns = engine.summarize(lists)
    for i, s in enumerate(SUMM):
        if not s["scores"]:
            assert gt[i] == 255 and ns[i] == 0
            continue
        assert ns[i] == len(s["scores"])
        assert qs[i] == s["QS"], i                      # numpy's pairwise summation order reproduced: bit-equal
        assert gs[i] == s["GS"]
        assert ["0/0", "0/1", "1/1"][gt[i]] == s["GT"], i
        assert abs(gq[i] - s["GQ"]) <= 1e-3
