#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-sm}
timeout 600 python bench.py --n-sv 12500 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_12k.json 2> gpurun_out/${T}_12k.err; echo "rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_12k.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()})
print("e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 2), d["e2e"]["single_blocking_call"], "wall resident", d.get("wall_ms_per_step_resident"))
PY
