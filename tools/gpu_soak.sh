#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/soak_parity.py --recipe simple --n-sv 4000 --k2-mode 2 > gpurun_out/r02_soak_simple.json 2> gpurun_out/r02_soak_simple.err; echo "simple rc=$?"; cat gpurun_out/r02_soak_simple.json
python tools/soak_parity.py --recipe complex --n-sv 3500 --k2-mode 2 > gpurun_out/r02_soak_complex.json 2> gpurun_out/r02_soak_complex.err; echo "complex rc=$?"; cat gpurun_out/r02_soak_complex.json
python tools/soak_parity.py --recipe large --n-sv 64 --k2-mode 2 > gpurun_out/r02_soak_large.json 2> gpurun_out/r02_soak_large.err; echo "large rc=$?"; cat gpurun_out/r02_soak_large.json
