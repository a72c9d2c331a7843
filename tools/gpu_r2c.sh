#!/bin/bash
# default bench (config 5, 100k SVs) at N GPUs, both arms
cd "$(dirname "$0")/.."
N=${1:-1}; STEPS=${2:-10}; TAG=${3:-r2c}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
SECONDS=0; $LAUNCH bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
echo "rc=$? wall=${SECONDS}s"; tail -c 600 gpurun_out/${TAG}_n$N.err | head -c 400
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_n$N.json"))
print({k:d[k] for k in ["value","ms_per_step","n_gpus","scaling","output_checksum","phase_ms_per_step","workload_gen_s","partition"]})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["single_blocking_call"])
print("cpu", d["cpu_baseline"])
PY
