#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02j
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c5_n1.json 2> gpurun_out/${T}_bench_c5_n1.err; echo "c5 rc=$?"; tail -3 gpurun_out/${T}_bench_c5_n1.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_bench_c5_reference.json 2> gpurun_out/${T}_bench_c5_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02j_bench_c5_n1.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["output_checksum"][:12])
print({k: (round(v["ms"], 1), round(v["frac"], 3)) for k, v in d["kernels"].items()})
print({k: v for k, v in d["roofline"].items() if k not in ("note",)})
print(d["cpu_baseline"])
PY
