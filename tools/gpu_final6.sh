#!/bin/bash
# closing pass of round 2: GPU tests, smoke, the default bench line (both arms) with the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02k
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c5_n1.json 2> gpurun_out/${T}_bench_c5_n1.err; echo "c5 rc=$?"; tail -2 gpurun_out/${T}_bench_c5_n1.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_bench_c5_reference.json 2> gpurun_out/${T}_bench_c5_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02k_bench_c5_n1.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["output_checksum"][:12])
print({k: (round(v["ms"], 1), round(v["frac"], 3)) for k, v in d["kernels"].items()}, d["roofline"]["kernel"], d["roofline"]["traffic"], d["cpu_baseline"]["value"], d["cpu_baseline"]["gpu_scores_match_on_sample"])
r = json.load(open("gpurun_out/r02k_bench_c5_reference.json")); print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
