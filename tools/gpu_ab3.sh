#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ab3}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
for w in 3 4; do
  timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --opt k3_warp_classes=$w > gpurun_out/${T}_w$w.json 2> gpurun_out/${T}_w$w.err; echo "warp_classes=$w rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_w*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["output_checksum"][:12])
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 python tools/soak_parity.py --recipe simple --n-sv 1500 > gpurun_out/${T}_soak_simple.json 2> gpurun_out/${T}_soak_simple.err; echo "soak rc=$?"; cat gpurun_out/${T}_soak_simple.json; tail -3 gpurun_out/${T}_soak_simple.err
