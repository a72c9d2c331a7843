#!/bin/bash
# A/B on config 5: overlap of kernel 3 with the next wave (0/1) x classes given to the warp kernel (0/3/5)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ab2}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
for cfg in "1 3" "0 3" "1 0" "1 5" "0 0"; do
  set -- $cfg
  timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --opt overlap=$1 --opt k3_warp_classes=$2 > gpurun_out/${T}_ovl$1_w$2.json 2> gpurun_out/${T}_ovl$1_w$2.err; echo "overlap=$1 warp_classes=$2 rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_ovl*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["output_checksum"][:12])
    except Exception as e:
        print(f, "ERR", e)
PY
