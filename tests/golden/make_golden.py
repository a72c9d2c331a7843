"""Generate tests/golden/*.json from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/vapor_vali/Simple_function.pyx through oracle.reference_loader (plain-Python
import of the untouched file, or its Cython build under oracle/_ref) and records, for seeded inputs:
  * dotdata hit lists (count, order-independent coordinate checksum, head and tail of the list),
  * the [a, b] pairs of the three calcu_vapor_single_read_score_* modes,
  * sizes of the cleaned dot sets and the re-centring intercept,
  * result_organize_ins rows and gt_estimate_log_likelihood outputs for score vectors.
The fixtures travel to the GPU box, where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.batch_oracle import hit_checksum          # noqa: E402
from oracle.reference_loader import load_reference    # noqa: E402
from vapor_b200 import synth                          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = ["calcu_vapor_single_read_score_abs_dis_m1b", "calcu_vapor_single_read_score_within_10Perc_m1b",
         "calcu_vapor_single_read_score_directed_dis_m1b_redefine_diagnal"]


def _f(v):
    return float(v)


def dot_record(R, k, read, struct):
    d = R.dotdata(k, read, struct)
    a = np.array(d, dtype=np.int64).reshape(-1, 2)
    return {"n": len(d), "checksum": str(hit_checksum(a)), "head": [list(map(int, t)) for t in d[:40]],
            "tail": [list(map(int, t)) for t in d[-10:]]}


def case_record(R, name, read, ref, alt, miss, k):
    rec = {"name": name, "read": read, "ref": ref, "alt": alt, "miss": miss, "k": k}
    x = [read, miss, "q"]
    try:
        rec["dot_ref"] = dot_record(R, k, read, ref[miss:])
        rec["dot_alt"] = dot_record(R, k, read, alt[miss:])
        rec["dot_ref_upper"] = dot_record(R, k, read, ref.upper()[miss:])
    except KeyError as e:
        rec["error"] = "KeyError"
        return rec
    for m in MODES:
        p = getattr(R, m)(ref, alt, x, k)
        rec[m] = [_f(p[0]), _f(p[1])]
    d = R.dotdata(k, read, ref[miss:])
    if len(d) > 0:
        c = R.clean_dotdata_diagnal_and_anti_diagnal(d)
        rec["n_clean_a6"] = len(c)
        rec["n_clean_diag_m1b"] = len(R.clean_dotdata_diagnal_m1b(d)[0])
        rec["n_clean_anti_m1b"] = len(R.clean_dotdata_anti_diagnal_m1b(d)[0])
        if len(c) > 0:
            rec["intercept"] = _f(R.dis_to_diagnal_most_abundant_defined([list(t) for t in c]))
    return rec


def main():
    R = load_reference()
    assert R is not None, "reference not available"
    rng = np.random.default_rng(20261018)
    cases = []
    # simulated SV cases, every type, several k, both haplotypes, miss_bp > 0, a soft-masked stretch
    i = 0
    for st in synth.SV_TYPES:
        for k in (10, 20, 30, 40):
            for hap_alt in (False, True):
                ln = int(rng.integers(60, 900))
                case = synth.make_sv_case(rng, st, ln, genotype=1, k=k, micro_indel=(i % 3 == 0),
                                          lowercase_frac=0.25 if i % 4 == 1 else 0.0)
                hap = case.hap_alt if hap_alt else case.hap_ref
                want = case.read_window - (i % 3)
                reads, _ = synth.simulate_reads(rng, hap, np.array([i % 3]), np.array([min(len(hap) - 3, int(want * 1.12) + 60)]),
                                                np.array([want]), err=0.15 if k == 10 else 0.04)
                cases.append(case_record(R, f"{st}_k{k}_{'alt' if hap_alt else 'ref'}", reads.tobytes().decode(),
                                         case.ref_seq.tobytes().decode(), case.alt_seq.tobytes().decode(), i % 3, k))
                i += 1
    # hand-made edge cases
    s = synth.random_dna(rng, 400).tobytes().decode()
    pal = "ACGTACGTAC" + "GTACGTACGT"                       # 20-mer that is its own reverse complement
    edge = [
        ("palindrome_k4", "ACGTACGTAA", "ACGTTTACGT", "ACGTTTACGT", 0, 4),
        ("palindrome_k20", s[:100] + pal + s[100:200], s[:100] + pal + s[100:200], s[:150], 0, 20),
        ("n_run", s[:80] + "N" * 30 + s[110:220], s[:80] + "N" * 30 + s[110:220], s[:220], 0, 10),
        ("iupac", s[:50] + "R" + s[51:200], s[:50] + "Y" + s[51:200], s[:200], 0, 10),
        ("lower_struct", s[:300], s[:100] + s[100:200].lower() + s[200:300], s[:300].lower(), 0, 10),
        ("lower_read", s[:100] + s[100:200].lower() + s[200:300], s[:100] + s[100:200].lower() + s[200:300], s[:300], 0, 10),
        ("x_in_alt", s[:300], s[:300], s[:100] + "X" * 100 + s[100:300], 0, 10),
        ("short_read", "ACGTA", s[:100], s[:100], 0, 10),
        ("short_struct", s[:100], "ACGTA", "ACG", 0, 10),
        ("polyA", "A" * 120, "A" * 150, "A" * 90 + "C" * 60, 0, 10),
        ("tandem_repeat", ("ACGGTCA" * 40)[:260], ("ACGGTCA" * 50)[:300], s[:300], 0, 10),
        ("miss_beyond", s[:100], s[:120], s[:50], 70, 10),
        ("revcomp_read", synth.revcomp(np.frombuffer(s[:300].encode(), np.uint8)).tobytes().decode(), s[:300], s[:300], 0, 10),
        ("bad_read_char", s[:50] + "X" + s[51:120], s[:120], s[:120], 0, 10),
        ("identical", s[:350], s[:350], s[:350], 0, 10),
    ]
    for name, read, ref, alt, miss, k in edge:
        cases.append(case_record(R, name, read, ref, alt, miss, k))
    json.dump({"generator": "tests/golden/make_golden.py", "reference": R.__vapor_kind__, "cases": cases},
              open(os.path.join(HERE, "scoring_cases.json"), "w"))
    # per-SV summaries
    srng = np.random.default_rng(7)
    summ = []
    for n in list(range(1, 26)) + [40, 64, 100, 130, 200]:
        for rep in range(4):
            kind = rep % 4
            if kind == 0:
                sc = srng.uniform(-1.5, 1.0, n)
            elif kind == 1:
                sc = srng.uniform(0.0, 1.0, n)
            elif kind == 2:
                sc = -srng.uniform(0.0, 3.0, n)
            else:
                sc = srng.choice([0.004, 0.005, 0.0051, -0.003, 0.476190476, -0.909090909, 0.015, 0.0149], n)
            sc = [float(v) for v in sc]
            row = R.result_organize_ins(["key", sc])
            gt = R.gt_estimate_log_likelihood(row)
            summ.append({"scores": sc, "QS": _f(row[1]), "GS": _f(row[2]), "Rec": row[3], "GT": gt[0], "GQ": _f(gt[1])})
    summ.append({"scores": [], "row": R.result_organize_ins(["key", []])[1:]})
    json.dump({"generator": "tests/golden/make_golden.py", "reference": R.__vapor_kind__, "summaries": summ},
              open(os.path.join(HERE, "sv_summaries.json"), "w"))
    print("cases", len(cases), "summaries", len(summ))


if __name__ == "__main__":
    main()
