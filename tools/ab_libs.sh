#!/bin/bash
# A/B of library builds on ONE box: bench.py (fast, 2^27 cells) and the ensemble config for every .so given
for lib in "$@"; do
  echo "== $lib"
  HRWENO_B200_LIB=$PWD/$lib timeout 200 python bench.py --mode fast --single-mode --steps 10 --warmup 3 --log2-cells 27 --no-cpu-baseline --no-extra-configs 2>&1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 2^27 fast', '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'])
except Exception as e: print('FAILED', e)"
  HRWENO_B200_LIB=$PWD/$lib timeout 200 python tools/bench_configs.py --mode ${MODE:-fast} --only ${ONLY:-cfg5} 2>&1 | sed 's/ensemble 65536x4096 //'
done
