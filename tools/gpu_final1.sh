#!/bin/bash
# final measurement pass, 1 GPU: default bench (both arms), configs 2/3/4, tile-kernel line, k1/k1b ncu, budget experiment
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02f
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_c5_n1.json 2> gpurun_out/${T}_bench_c5_n1.err; echo "c5 rc=$?"
python bench.py --steps 10 --warmup 3 --hit-budget-gb 40 --no-cpu-baseline > gpurun_out/${T}_bench_c5_n1_budget40.json 2> /dev/null; echo "c5 budget40 rc=$?"
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_bench_c5_reference.json 2> gpurun_out/${T}_bench_c5_reference.err; echo "ref rc=$?"
for c in 2 3 4; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/${T}_bench_c${c}_n1.json 2> gpurun_out/${T}_bench_c${c}_n1.err; echo "c$c rc=$?"
done
python bench.py --config 2 --steps 5 --warmup 3 --k2-mode 0 --no-cpu-baseline > gpurun_out/${T}_bench_c2_n1_tile_kernel.json 2> /dev/null; echo "c2 tile rc=$?"
python bench.py --config 4 --steps 3 --warmup 2 --k2-mode 0 --no-cpu-baseline > gpurun_out/${T}_bench_c4_n1_tile_kernel.json 2> /dev/null; echo "c4 tile rc=$?"
python tools/perf_k.py 10 2000 > gpurun_out/${T}_plain_perf.log 2>&1
for RX in k1_pack_kmers k1b_build_tables; do
  ncu --set full --clock-control none --import-source on -k regex:$RX -s 1 -c 1 -f -o gpurun_out/r02_full_$RX python tools/perf_k.py 10 2000 > gpurun_out/${T}_ncu_$RX.log 2>&1; echo "ncu $RX rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02f_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), d.get("phase_ms_per_step"), (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("gpu_scores_match_on_sample"))
    except Exception as e:
        print(f, "ERR", e)
PY
