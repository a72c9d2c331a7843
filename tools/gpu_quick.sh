#!/bin/bash
# The inner loop of kernel work, one gpurun call: GPU tests, the default bench line (its output_checksum must not move), optionally an
# ncu --set full capture of the launches matching a kernel regex on tools/perf_k.py 10 2000 (read here with tools/ncu_summary.py and
# tools/ncu_by_line.py).      gpu_quick.sh <tag> [kernel-regex] [launches-to-skip] [launches-to-capture]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-quick}; RX=$2; SKIP=${3:-0}; CNT=${4:-5}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c5.json 2> gpurun_out/${T}_c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/${T}_c5.err
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_c5.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["output_checksum"][:12])
print({k: (round(v["ms"], 1), round(v["frac"], 3)) for k, v in d["kernels"].items()})
PY
if [ -n "$RX" ]; then
  python tools/perf_k.py 10 2000 > gpurun_out/${T}_plain.log 2>&1 && cat gpurun_out/${T}_plain.log &&
  ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -f -o gpurun_out/prof_${T} python tools/perf_k.py 10 2000 > gpurun_out/${T}_ncu.log 2>&1
  tail -2 gpurun_out/${T}_ncu.log
fi
