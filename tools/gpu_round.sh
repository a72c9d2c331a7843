# one GPU call: tests, bench (both arms), ncu launch list, ncu full capture of the tile kernel
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_10k.json 2> gpurun_out/bench_10k.err; tail -c 600 gpurun_out/bench_10k.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --n-sv 1000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_list.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --n-sv 1000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
python bench.py --n-sv 400 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k2_tile -s 1 -c 1 -o gpurun_out/prof_k2 -f \
    python bench.py --n-sv 400 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench_10k.json
