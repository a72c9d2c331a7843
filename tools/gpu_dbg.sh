#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/dbg_hits.py 400 2>&1 | tee gpurun_out/dbg_hits.log
for tool in racecheck initcheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 8 python tools/dbg_hits.py 24 > gpurun_out/dbg_$tool.log 2>&1; echo "$tool rc=$?"; grep -v "^k2_mode" gpurun_out/dbg_$tool.log | head -40
done
