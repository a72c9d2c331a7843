#!/bin/bash
# warp-per-task kernel 3: GPU tests, A/B bench of k3_mode on config 5 (checksums must agree), small soak
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2d}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${T}_tests.log
for m in 1 0; do
  timeout 600 python bench.py --steps 5 --warmup 3 --k3-mode $m --no-cpu-baseline > gpurun_out/${T}_c5_k3mode$m.json 2> gpurun_out/${T}_c5_k3mode$m.err; echo "c5 k3_mode=$m rc=$?"
done
timeout 600 python tools/soak_parity.py --recipe simple --n-sv 1500 > gpurun_out/${T}_soak_simple.json 2> gpurun_out/${T}_soak_simple.err; echo "soak rc=$?"; cat gpurun_out/${T}_soak_simple.json; tail -3 gpurun_out/${T}_soak_simple.err
timeout 600 python tools/soak_parity.py --recipe complex --n-sv 1200 > gpurun_out/${T}_soak_complex.json 2> gpurun_out/${T}_soak_complex.err; echo "soak rc=$?"; cat gpurun_out/${T}_soak_complex.json; tail -3 gpurun_out/${T}_soak_complex.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_c5_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), d.get("phase_ms_per_step"), d["output_checksum"])
    except Exception as e:
        print(f, "ERR", e)
PY
