#!/bin/bash
# cfg3 roofline fraction of one library at several (log2 cells, steps) combinations on ONE box
lib=$1; shift
for combo in "$@"; do
  l2=${combo%%:*}; st=${combo##*:}
  HRWENO_B200_LIB=$PWD/$lib timeout 300 python bench.py --mode fast --single-mode --steps $st --warmup 3 --log2-cells $l2 --no-cpu-baseline --no-extra-configs 2>&1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib'.split('/')[-1].ljust(22), '2^$l2 steps $st', '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'stage_ms %.4f'%d['roofline']['avg_launch_ms'], 'e2e %.3e'%d['e2e']['value'], 'sm %s'%d['clocks']['sm_mhz'], 'W %s'%d['clocks']['power_w_max'])
except Exception as e: print('$lib', 'FAILED', e)"
done
