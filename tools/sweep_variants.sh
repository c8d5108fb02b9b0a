#!/bin/bash
# bench every tuning variant of the library (hr-weno_b200/lib/variants/*.so) in both modes; prints one line each
for lib in hr-weno_b200/lib/libhrweno_b200.so hr-weno_b200/lib/variants/*.so; do
  for m in strict fast; do
    HRWENO_B200_LIB=$PWD/$lib timeout 200 python bench.py --mode $m --steps 5 --warmup 3 --log2-cells 26 --no-cpu-baseline --no-extra-configs 2>&1 | \
      python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib'.split('/')[-1], d['config']['mode'], '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'])
except Exception as e: print('$lib', '$m', 'FAILED', e)"
  done
  HRWENO_B200_LIB=$PWD/$lib timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
done
