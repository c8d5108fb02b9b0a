#!/bin/bash
# compute-sanitizer over the stage kernels at small sizes; logs under gpurun_out/ (summaries are copied to profiles/)
out=${1:-gpurun_out}
for tool in memcheck racecheck synccheck; do
  for part in 1d small 2d general pipeline mgpu; do
    echo "== $tool $part"
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_driver.py $part > $out/sanitize_${tool}_${part}.log 2>&1
    echo "rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/sanitize_${tool}_${part}.log | tail -1)"
  done
done
