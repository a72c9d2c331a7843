for mb in 5 6 8 1; do
  VAPOR_NVCC_EXTRA="-DK3_MINB=$mb" python -c "from vapor_b200 import _build; _build.build_native(force=True)"
  python bench.py --n-sv 2000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k3mb$mb.json 2> /dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_k3mb$mb.json'))
print('k3 minb $mb', d['phase_ms_per_step'])
PY
done
