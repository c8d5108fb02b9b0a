#!/bin/bash
# A/B of 1D stage-kernel library variants on ONE box: cfg3 (fast, 2^27 cells, 10 steps) twice per library, then the tiled
# fast-mode oracle parity tests for each variant.  usage: tools/ab_k2.sh lib1.so lib2.so ...
for rep in 1 2; do
for lib in "$@"; do
  HRWENO_B200_LIB=$PWD/$lib timeout 200 python bench.py --mode fast --single-mode --steps 10 --warmup 3 --log2-cells 27 --no-cpu-baseline --no-extra-configs 2>&1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib'.split('/')[-1].ljust(24), 'cfg3 2^27 fast', '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'], 'parity %.2e'%d['parity_check']['max_normwise'], 'sm %s'%d['clocks']['sm_mhz'])
except Exception as e: print('$lib', 'FAILED', e)"
done
done
if [ -n "$PARITY" ]; then
for lib in "$@"; do
  echo "== parity $lib"
  HRWENO_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_fast_tiled.py -x -q -m gpu 2>&1 | tail -2
done
fi
