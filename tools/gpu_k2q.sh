#!/bin/bash
# quick loop for the join kernel: GPU tests, config-5 bench (checksum must stay 0725a883...), ncu capture of its launches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k2q}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c5.json 2> gpurun_out/${T}_c5.err; echo "c5 rc=$?"; tail -3 gpurun_out/${T}_c5.err
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_c5.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), d.get("phase_ms_per_step"), d["output_checksum"][:12], "evaluated", d["roofline"].get("cells_evaluated"))
PY
if [ "$2" != "noncu" ]; then
python tools/perf_k.py 10 2000 > gpurun_out/plain_k2q.log 2>&1 && cat gpurun_out/plain_k2q.log &&
ncu --set full --clock-control none --import-source on -k regex:k2_join -s 6 -c 3 -o gpurun_out/prof_${T} -f \
    python tools/perf_k.py 10 2000 > gpurun_out/ncu_k2q.log 2>&1
tail -2 gpurun_out/ncu_k2q.log
fi
