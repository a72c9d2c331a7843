#!/bin/bash
# final measurement pass of round 2 (second half), 1 GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02g
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_c5_n1.json 2> gpurun_out/${T}_bench_c5_n1.err; echo "c5 rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt overlap=1 > gpurun_out/${T}_bench_c5_n1_overlap.json 2> /dev/null; echo "c5 overlap rc=$?"
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_bench_c5_reference.json 2> gpurun_out/${T}_bench_c5_reference.err; echo "ref rc=$?"
for c in 2 3 4; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/${T}_bench_c${c}_n1.json 2> gpurun_out/${T}_bench_c${c}_n1.err; echo "c$c rc=$?"
done
python bench.py --config 2 --steps 5 --warmup 3 --k2-mode 0 --no-cpu-baseline > gpurun_out/${T}_bench_c2_n1_tile_kernel.json 2> /dev/null; echo "c2 tile rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_nsv25000.csv python bench.py --n-sv 25000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/perf_k.py 10 2000 > gpurun_out/${T}_plain_perf.log 2>&1; cat gpurun_out/${T}_plain_perf.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_pack|k1b_build|k2_join|k3" -s 20 -c 10 -f -o gpurun_out/${T}_full python tools/perf_k.py 10 2000 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${T}_ncu_full.log
python tools/soak_parity.py --recipe simple --n-sv 2500 --k2-mode 2 > gpurun_out/${T}_soak_simple.json 2> gpurun_out/${T}_soak_simple.err; echo "soak simple rc=$?"; cat gpurun_out/${T}_soak_simple.json
python tools/soak_parity.py --recipe complex --n-sv 2000 --k2-mode 2 > gpurun_out/${T}_soak_complex.json 2> gpurun_out/${T}_soak_complex.err; echo "soak complex rc=$?"; cat gpurun_out/${T}_soak_complex.json
python tools/soak_parity.py --recipe large --n-sv 48 --k2-mode 2 > gpurun_out/${T}_soak_large.json 2> gpurun_out/${T}_soak_large.err; echo "soak large rc=$?"; cat gpurun_out/${T}_soak_large.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02g_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in (d.get("phase_ms_per_step") or {}).items()}, (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("gpu_scores_match_on_sample"), (d.get("output_checksum") or "")[:12])
    except Exception as e:
        print(f, "ERR", e)
PY
