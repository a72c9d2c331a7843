#!/bin/bash
# closing pass with the final build: GPU tests, default bench (both arms), smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02i
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c5_n1.json 2> gpurun_out/${T}_bench_c5_n1.err; echo "c5 rc=$?"
for c in 2 3 4; do
  python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_c${c}_n1.json 2> gpurun_out/${T}_bench_c${c}_n1.err; echo "c$c rc=$?"
done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02i_bench_*.json")):
    d = json.load(open(f))
    print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["output_checksum"][:12], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), d["roofline"]["traffic"])
PY
