#!/bin/bash
# A/B of 2D stage-kernel library variants on ONE box: cfg4 (16384^2, WENO5+mstvd) in both modes, then the 2D parity tests.
# usage: [PARITY=1] tools/ab_k3.sh lib1.so lib2.so ...
for lib in "$@"; do
  for m in fast strict; do
    echo -n "$(basename $lib) $m: "
    HRWENO_B200_LIB=$PWD/$lib timeout 300 python tools/bench_configs.py --mode $m --only cfg4 2>&1 | tail -1
  done
done
if [ -n "$PARITY" ]; then
for lib in "$@"; do
  echo "== parity $lib"
  HRWENO_B200_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_tiled.py tests/test_zzz_gpu_reference_source.py -x -q -m gpu -k "2d or example2 or config2 or ms" 2>&1 | tail -3
done
fi
