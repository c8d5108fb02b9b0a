#!/bin/bash
# gpu_retry.sh TIMEOUT 'command' : gpurun with retries while the pod has no free slot (exit code 3 / "transient")
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpu_retry_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpu_retry_last.log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
tail -n 80 /tmp/gpu_retry_last.log
exit $rc
