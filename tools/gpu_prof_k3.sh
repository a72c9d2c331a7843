python bench.py --n-sv 1000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_k3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k3_score -s 3 -c 3 -o gpurun_out/prof_k3 -f \
    python bench.py --n-sv 1000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_k3.log 2>&1
tail -2 gpurun_out/ncu_k3.log
