#!/bin/bash
# smoke of the new bench: every config, small sizes, 1 GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "2 2000" "3 1000" "4 40" "5 4000"; do
  set -- $spec
  timeout 900 python bench.py --config $1 --n-sv $2 --steps 3 --warmup 3 --ref-seconds 1 > gpurun_out/r2b_c$1.json 2> gpurun_out/r2b_c$1.err
  echo "config $1 rc=$?"; tail -3 gpurun_out/r2b_c$1.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_c$1.json"))
    print({k:d[k] for k in ["value","ms_per_step","output_checksum","phase_ms_per_step","workload_gen_s"]})
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["single_blocking_call"])
    print("roofline", {k:v for k,v in d["roofline"].items() if k in ("kernel","bound","achieved","peak","frac","share_of_step")})
    print("cpu", d["cpu_baseline"])
except Exception as e: print("no json", e)
PY
done
