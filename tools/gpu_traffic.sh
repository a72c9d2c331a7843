#!/bin/bash
# GPU tests with the final build, then DRAM traffic per launch of the wave kernels at the benchmark's wave size
# (a wave of the 100k-SV list = 6250 SVs; 25k SVs = 4 such waves)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02i_tests.log
timeout 900 ncu --set full --clock-control none -k regex:"k1_pack|k1b_build|k2_join|k3" -s 0 -c 10 -f -o gpurun_out/r02h_wave_full python bench.py --n-sv 25000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02h_wave_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02h_wave_ncu.log
ls -la gpurun_out/r02h_wave_full.ncu-rep
