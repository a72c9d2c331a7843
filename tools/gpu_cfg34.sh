#!/bin/bash
# GPU tests, then the bench lines of configs 3 and 4 (hashed words: k > 15) and a large-window soak
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-c34}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
for c in 3 4; do
  python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_c${c}_n1.json 2> gpurun_out/${T}_bench_c${c}_n1.err; echo "c$c rc=$?"
done
python tools/soak_parity.py --recipe large --n-sv 48 --k2-mode 2 > gpurun_out/${T}_soak_large.json 2> gpurun_out/${T}_soak_large.err; echo "soak large rc=$?"; cut -c1-400 gpurun_out/${T}_soak_large.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    d = json.load(open(f))
    print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["output_checksum"][:12])
PY
