for lib in hr-weno_b200/lib/libhrweno_b200.so hr-weno_b200/lib/variants/*.so; do
  HRWENO_B200_LIB=$PWD/$lib timeout 200 python bench.py --mode fast --single-mode --steps 10 --warmup 3 --log2-cells 27 --no-cpu-baseline --no-extra-configs 2>&1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib'.split('/')[-1], d['config']['mode'], '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'])
except Exception as e: print('$lib', 'FAILED', e)"
done
