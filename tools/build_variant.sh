#!/bin/bash
# build_variant.sh NAME "EXTRA flags" [FAST_MINB] : library variant hr-weno_b200/lib/variants/NAME.so whose 1D stage objects are
# compiled with the given macros (the other objects are taken from the main build directory)
set -e
cd "$(dirname "$0")/../hr-weno_b200/csrc"
name=$1; extra=$2; minb=${3:-4}
rm -rf build_$name; cp -rp build build_$name; rm -f build_$name/fv1d_inst_*.o
mkdir -p ../lib/variants
make -j8 BUILD=build_$name LIB=../lib/variants/$name.so EXTRA="$extra" FAST_MINB=$minb 2>&1 | grep -E "error|spill stores|warning" | grep -v " 0 bytes spill" || true
ls -la ../lib/variants/$name.so
