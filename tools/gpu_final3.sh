#!/bin/bash
# last pass with the final build: default bench (both arms), launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02h
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c5_n1.json 2> gpurun_out/${T}_bench_c5_n1.err; echo "c5 rc=$?"
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_bench_c5_reference.json 2> gpurun_out/${T}_bench_c5_reference.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_nsv25000.csv python bench.py --n-sv 25000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02h_bench_c5_n1.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["cpu_baseline"]["value"], d["cpu_baseline"]["gpu_scores_match_on_sample"], d["output_checksum"][:12], d["e2e"]["pipeline_stage_ms_rank0"])
PY
