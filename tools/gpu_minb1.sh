for mb in 4 6 8; do
  VAPOR_NVCC_EXTRA="-DK1_MINB=$mb" python -c "from vapor_b200 import _build; _build.build_native(force=True)"
  for k in 10 40; do python tools/perf_k.py $k 2>&1 | tail -1 | cut -c1-60 | sed "s/^/k1 minb $mb /"; done
done
