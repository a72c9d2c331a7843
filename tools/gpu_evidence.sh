#!/bin/bash
# round-2 evidence pass: integer peaks, ncu launch list of the default bench command, ncu --set full of every kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
python tools/int_peaks.py > gpurun_out/${TAG}_int_peaks.json 2> gpurun_out/${TAG}_int_peaks.err; cat gpurun_out/${TAG}_int_peaks.json
# launch list of the bench command (cold-cache, serialised: compare SHARES)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_plain_bench.json 2> gpurun_out/${TAG}_plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_bench_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/${TAG}_launches_bench_default.csv
# full captures on a short resident run
python tools/perf_k.py 10 2000 > gpurun_out/${TAG}_plain_perf.log 2>&1 || { echo "plain perf failed"; exit 1; }
for RX in k1_pack_kmers k1b_build_tables k2_join_match k3_score_reads; do
  ncu --set full --clock-control none --import-source on -k regex:$RX -s 3 -c 3 -f -o gpurun_out/${TAG}_full_$RX \
      python tools/perf_k.py 10 2000 > gpurun_out/${TAG}_ncu_$RX.log 2>&1
  echo "ncu $RX rc=$?"
done
