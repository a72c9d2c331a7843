set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for v in ${VARIANTS:-0 1 2}; do
  python bench.py --n-sv ${NSV:-2000} --steps 3 --warmup 3 --no-cpu-baseline --tile-variant $v > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_v$v.json'))
print('variant $v', 'cells/s %.3e'%d['cells_per_sec'], 'reads/s %.0f'%d['value'], d['phase_ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d.get('tile_padding'))
PY
done
