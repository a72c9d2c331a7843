#!/bin/bash
# quick loop for the warp kernel 3: config-5 bench (checksum must stay 0725a883...), then an ncu capture of its five launches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k3wq}
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c5.json 2> gpurun_out/${T}_c5.err; echo "c5 rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_c5.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), d.get("phase_ms_per_step"), d["output_checksum"][:12])
PY
if [ "$2" != "noncu" ]; then
python tools/perf_k.py 10 2000 > gpurun_out/plain_k3w.log 2>&1 && cat gpurun_out/plain_k3w.log &&
ncu --set full --clock-control none --import-source on -k regex:k3w_score -s 10 -c 5 -o gpurun_out/prof_${T} -f \
    python tools/perf_k.py 10 2000 > gpurun_out/ncu_k3w.log 2>&1
tail -2 gpurun_out/ncu_k3w.log
fi
