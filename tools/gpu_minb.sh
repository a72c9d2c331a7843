for mb in 5 6 4; do
  VAPOR_NVCC_EXTRA="-DK2_MINB=$mb" python -c "from vapor_b200 import _build; _build.build_native(force=True)"
  python bench.py --n-sv 2000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mb$mb.json 2> gpurun_out/bench_mb$mb.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_mb$mb.json'))
print('minb $mb', 'cells/s %.3e'%d['cells_per_sec'], d['phase_ms_per_step'])
PY
done
