#!/bin/bash
# gpu_retry.sh [-g N] TIMEOUT 'command' : gpurun with retries while the pod has no free slot (exit code 3 / "transient")
g=1
if [ "$1" = "-g" ]; then g=$2; shift 2; fi
t=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --gpus $g --timeout $t -- "$@" > /tmp/gpu_retry_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpu_retry_last.log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
tail -n 80 /tmp/gpu_retry_last.log
exit $rc
