#!/usr/bin/env python
"""k selection on repeat-rich windows: the reference's window_size_refine (vapor_vali/Simple_function.pyx:2030-2046,
unmodified, with numpy's random state seeded before every call -- its KMeans / scipy kmeans are otherwise unseeded)
against vapor_b200's refine_many on the same windows with the same seeding.  BUILD CONTAINER ONLY (needs /root/reference).

    python tools/kselect_agreement.py [--n 2000] [--out profiles/r02_kselect_agreement.json]

Windows: random sequence with planted tandem arrays (3-12 copies of a 15-600 bp unit, 0-10 % divergence per copy),
inverted repeats and interspersed copies -- the cases in which fewer than 40 % of the self-plot dots lie on the diagonal.
The engine here is the CPU oracle stand-in (tests/oracle_engine.py); the GPU suite checks that the CUDA self-plot
counters equal the oracle's (tests/test_gpu_parity.py::test_selfplot_qc_counts)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def make_window(rng):
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    rnd = lambda n: acgt[rng.integers(0, 4, size=n, dtype=np.uint8)]
    def mutate(a, rate):
        a = a.copy()
        m = rng.random(len(a)) < rate
        a[m] = acgt[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
        return a
    kind = rng.choice(["tandem", "tandem", "tandem", "inverted", "interspersed", "mixed"])
    parts = [rnd(int(rng.integers(50, 500)))]
    if kind in ("tandem", "mixed"):
        unit = rnd(int(rng.choice([15, 23, 40, 77, 150, 300, 600])))
        copies = int(rng.integers(3, 13))
        div = float(rng.choice([0.0, 0.01, 0.03, 0.06, 0.1]))
        parts += [mutate(unit, div) for _ in range(copies)]
    if kind in ("inverted", "mixed"):
        arm = rnd(int(rng.integers(100, 900)))
        comp = np.zeros(256, np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
        parts += [arm, rnd(int(rng.integers(10, 300))), comp[arm[::-1]]]
    if kind == "interspersed":
        el = rnd(int(rng.integers(80, 400)))
        for _ in range(int(rng.integers(3, 9))):
            parts += [mutate(el, 0.03), rnd(int(rng.integers(20, 200)))]
    parts.append(rnd(int(rng.integers(50, 500))))
    return np.concatenate(parts).tobytes().decode(), str(kind)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--out", default="")
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--first", type=int, default=0)
    a = ap.parse_args()
    if a.procs > 1:
        import subprocess
        per = (a.n + a.procs - 1) // a.procs
        ps = [subprocess.Popen([sys.executable, __file__, "--n", str(min(per, a.n - i * per)), "--first", str(i * per)], stdout=subprocess.PIPE, text=True)
              for i in range(a.procs) if a.n - i * per > 0]
        tot = None
        for p in ps:
            d = json.loads(p.communicate()[0].strip().splitlines()[-1])
            if tot is None:
                tot = d
                continue
            for k, v in d.items():
                if isinstance(v, int):
                    tot[k] += v
                elif isinstance(v, dict):
                    for kk, vv in v.items():
                        if isinstance(vv, list):
                            cur = tot[k].setdefault(kk, [0, 0]); cur[0] += vv[0]; cur[1] += vv[1]
                        else:
                            tot[k][kk] = tot[k].get(kk, 0) + vv
                elif k == "seconds":
                    tot[k] = max(tot[k], v)
        tot["agreement"] = tot["agree"] / max(1, tot["windows"])
        tot["agreement_bounding_box_variant"] = tot["bounding_box_agree"] / max(1, tot["windows"])
        print(json.dumps(tot))
        if a.out:
            json.dump(tot, open(a.out, "w"), indent=1)
        return
    import scipy
    for n_ in ("std", "mean", "array", "sqrt"):            # SciPy dropped the numpy aliases the reference calls (Simple_function.pyx:878)
        if not hasattr(scipy, n_):
            setattr(scipy, n_, getattr(np, n_))
    from oracle.reference_loader import load_reference
    from oracle_engine import OracleEngine
    from vapor_b200 import Simple_function as SF
    ref = load_reference()
    sess = SF.Session(engine=OracleEngine())
    stats = {"windows": 0, "agree": 0, "ref_error": 0, "xmeans_branch": 0, "agree_in_branch": 0, "k_ref": {}, "k_ours": {},
             "bounding_box_agree": 0, "by_kind": {}}
    import warnings
    warnings.simplefilter("ignore")
    t0 = time.time()
    for i in range(a.first, a.first + a.n):
        seq, kind = make_window(np.random.default_rng([20261018, i]))
        np.random.seed(i)
        try:
            kr = ref.window_size_refine(seq)[0]
        except Exception as e:                              # noqa: BLE001
            kr = "raise:" + type(e).__name__
            stats["ref_error"] += 1
        np.random.seed(i)
        before = sess.stats["gpu_calls"]
        ko = sess.refine_many([SF.RefineRequest(seq)])[0][0]
        in_branch = sess.stats["gpu_calls"] - before > len({10, 20, 30, 40} & set(range(10, (ko if isinstance(ko, int) else 40) + 1, 10)))
        SF.XMEANS_ON_HOST = False
        kb = sess.refine_many([SF.RefineRequest(seq)])[0][0]
        SF.XMEANS_ON_HOST = True
        stats["windows"] += 1
        stats["agree"] += int(kr == ko)
        stats["bounding_box_agree"] += int(kr == kb)
        stats["xmeans_branch"] += int(in_branch)
        stats["agree_in_branch"] += int(in_branch and kr == ko)
        stats["k_ref"][str(kr)] = stats["k_ref"].get(str(kr), 0) + 1
        stats["k_ours"][str(ko)] = stats["k_ours"].get(str(ko), 0) + 1
        bk = stats["by_kind"].setdefault(kind, [0, 0])
        bk[0] += 1; bk[1] += int(kr == ko)
    stats["seconds"] = round(time.time() - t0, 1)
    stats["agreement"] = stats["agree"] / max(1, stats["windows"])
    stats["agreement_bounding_box_variant"] = stats["bounding_box_agree"] / max(1, stats["windows"])
    stats["note"] = ("same windows, numpy random state seeded identically before each call on both sides; 'bounding_box' = round 1's "
                     "single-box approximation (VAPOR_XMEANS=0)")
    print(json.dumps(stats))
    if a.out:
        json.dump(stats, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
