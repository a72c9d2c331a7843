python bench.py --n-sv 1000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_list.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -s 6 -c 12 --csv --log-file gpurun_out/launches2.csv \
    python bench.py --n-sv 1000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
