import sys, json
sys.path.insert(0, '.')
from vapor_b200 import synth
from vapor_b200.engine import Engine
k = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
w = synth.make_workload(n, seed=20261019, size_range=(50, 5000), reads_per_sv=20, k_choices=(k,), err=0.15, workers=12)
with Engine(0) as e:
    e.upload(w.batch)
    for _ in range(3):
        e.run()
    t = e.timings()
print(json.dumps({"k": k, "pack_ms": round(t["pack_ms"], 2), "tile_ms": round(t["tile_ms"], 2), "score_ms": round(t["score_ms"], 2),
                  "hits": t["hits"], "cells": t["cells"], "cells_per_s": t["cells"] / (t["tile_ms"] * 1e-3)}))
