set -x
V=${1:-1}
python bench.py --n-sv 400 --steps 1 --warmup 1 --no-cpu-baseline --tile-variant $V > gpurun_out/plain_v$V.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k2_tile -s 1 -c 1 -o gpurun_out/prof_k2_v$V -f \
    python bench.py --n-sv 400 --steps 1 --warmup 1 --no-cpu-baseline --tile-variant $V > gpurun_out/ncu_v$V.log 2>&1
tail -3 gpurun_out/ncu_v$V.log
