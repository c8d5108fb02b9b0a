for lib in "$@"; do
  echo "== $lib"
  HRWENO_B200_LIB=$PWD/$lib timeout 200 python bench.py --mode strict --single-mode --steps 5 --warmup 3 --log2-cells 27 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 2^27 strict %.3e frac %.3f' % (d['value'], d['roofline']['frac']))"
  HRWENO_B200_LIB=$PWD/$lib timeout 200 python tools/bench_configs.py --mode strict --only cfg5 2>&1 | grep rktvd3 | sed 's/ensemble 65536x4096 //'
done
