#!/bin/bash
# new kernel 1 (packed windows) + join kernel with shared plot metadata: tests, variants, bench, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k12}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
bash tools/gpu_variants.sh 4000 k1m6 it8 it16 2>&1 | tee gpurun_out/${T}_variants.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c5.json 2> gpurun_out/${T}_c5.err; echo "c5 rc=$?"; tail -3 gpurun_out/${T}_c5.err
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_c5.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), d.get("phase_ms_per_step"), d["output_checksum"][:12], "evaluated", d["roofline"].get("cells_evaluated"))
PY
python tools/perf_k.py 10 2000 > gpurun_out/plain_k2q.log 2>&1 && cat gpurun_out/plain_k2q.log &&
ncu --set full --clock-control none --import-source on -k regex:"k2_join|k1_pack" -s 8 -c 4 -o gpurun_out/prof_${T} -f \
    python tools/perf_k.py 10 2000 > gpurun_out/ncu_k2q.log 2>&1
tail -2 gpurun_out/ncu_k2q.log
