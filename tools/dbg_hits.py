import sys, json
sys.path.insert(0, '.')
from vapor_b200 import synth
from vapor_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
w = synth.make_workload(n, seed=20261019, size_range=(50, 5000), reads_per_sv=20, k_choices=(10,), err=0.15, workers=4)
with Engine(0) as e:
    for mode in (0, 1, 1, 0, 1):
        e.set_option("k2_mode", mode)
        e.upload(w.batch)
        e.run()
        t = e.timings()
        print("k2_mode", mode, "hits", t["hits"], "overflow_plots", t["n_overflow_plots"], flush=True)
