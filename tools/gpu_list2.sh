#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-l2}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c5.json 2> gpurun_out/${T}_c5.err; echo "c5 rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_c5.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["output_checksum"][:12])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --n-sv 25000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1; echo "ncu list rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${T}_launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[:75]:
    name = r[4].split("(")[0]; grid = r[8]; blk = r[7]
    key = (name, blk)
    agg[key][0] += 1; agg[key][1] += float(r[-1]) / 1e6
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(k, v)
# first wave in order
for r in rows[:22]: print(r[4].split("(")[0][:40], r[7], r[8], round(float(r[-1]) / 1e3, 1), "us")
PY
