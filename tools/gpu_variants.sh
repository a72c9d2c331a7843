#!/bin/bash
# time perf_k.py with each experimental build: gpu_variants.sh <n_sv> name1 name2 ...
cd "$(dirname "$0")/.."
N=$1; shift
echo -n "default: "; python tools/perf_k.py 10 $N
for v in "$@"; do echo -n "$v: "; VAPOR_B200_LIB=$PWD/vapor_b200/csrc/libvapor_b200_$v.so python tools/perf_k.py 10 $N; done
