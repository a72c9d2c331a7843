#!/bin/bash
# ncu --set full captures of named kernels on a short resident run (tools/perf_k.py): gpu_prof2.sh <n_sv> <tag> <regex> [<regex> ...]
cd "$(dirname "$0")/.."
N=${1:-1000}; TAG=${2:-r02}; shift 2
mkdir -p gpurun_out
python tools/perf_k.py 10 $N > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log
for RX in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$RX -s 3 -c 3 -f -o gpurun_out/${TAG}_$RX \
      python tools/perf_k.py 10 $N > gpurun_out/${TAG}_ncu_$RX.log 2>&1
  echo "ncu $RX rc=$?"; tail -2 gpurun_out/${TAG}_ncu_$RX.log
done
