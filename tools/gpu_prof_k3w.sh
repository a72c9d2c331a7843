#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/perf_k.py 10 2000 > gpurun_out/plain_k3w.log 2>&1 && cat gpurun_out/plain_k3w.log &&
ncu --set full --clock-control none --import-source on -k regex:k3w_score -s 10 -c 5 -o gpurun_out/prof_k3w -f \
    python tools/perf_k.py 10 2000 > gpurun_out/ncu_k3w.log 2>&1
tail -2 gpurun_out/ncu_k3w.log
