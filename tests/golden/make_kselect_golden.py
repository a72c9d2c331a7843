"""Golden vectors for k selection on repeat-rich windows (BUILD CONTAINER ONLY: runs the unmodified reference).

    python tests/golden/make_kselect_golden.py      -> tests/golden/kselect_cases.json

For 48 windows with planted tandem arrays / inverted / interspersed repeats (tools/kselect_agreement.make_window) the
reference's window_size_refine (vapor_vali/Simple_function.pyx:2030-2046) is called with numpy's random state seeded
(its KMeans / scipy kmeans are unseeded); the window, the seed and the k it returned are stored.  Windows that reach the
X-means branch (10-50 % of the dots below the diagonal and at most 40 % on it) are kept preferentially."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def main():
    import warnings
    warnings.simplefilter("ignore")
    import scipy
    for n_ in ("std", "mean", "array", "sqrt"):
        if not hasattr(scipy, n_):
            setattr(scipy, n_, getattr(np, n_))
    from kselect_agreement import make_window
    from oracle.reference_loader import load_reference
    from oracle import vapor_oracle as O
    ref = load_reference()
    cases, n_branch = [], 0
    i = 0
    while len(cases) < 48:
        seq, kind = make_window(np.random.default_rng([77, i]))
        i += 1
        if len(seq) > 2600:
            continue
        d = O.dotdata(10, seq, seq)
        low = int((d[:, 0] > d[:, 1]).sum()); diag = int((d[:, 0] == d[:, 1]).sum())
        branch = 0.1 < low / len(d) < 0.5 and diag / len(d) <= 0.4
        if not branch and len(cases) - n_branch >= 12:
            continue
        np.random.seed(1000 + i)
        k = ref.window_size_refine(seq)[0]
        cases.append({"seq": seq, "kind": kind, "seed": 1000 + i, "k": k, "xmeans_branch_at_k10": bool(branch)})
        n_branch += int(branch)
    json.dump({"generator": "tests/golden/make_kselect_golden.py", "reference": "vapor_vali/Simple_function.pyx window_size_refine",
               "cases": cases}, open(os.path.join(HERE, "kselect_cases.json"), "w"))
    print(len(cases), "cases,", n_branch, "in the X-means branch;", {k: sum(1 for c in cases if c["k"] == k) for k in (10, 20, 30, 40)})


if __name__ == "__main__":
    main()
