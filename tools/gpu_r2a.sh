#!/bin/bash
# round 2, first GPU pass: parity of both kernel-2 variants, then a first look at the join kernel's speed
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2a_tests.log
tail -15 gpurun_out/r2a_tests.log
for mode in 1 0; do
  timeout 600 python bench.py --steps 5 --warmup 3 --n-sv 2000 --k2-mode $mode --no-cpu-baseline > gpurun_out/r2a_bench_mode$mode.json 2> gpurun_out/r2a_bench_mode$mode.err
  echo "bench mode $mode rc=$?"; tail -c 1500 gpurun_out/r2a_bench_mode$mode.json
done
