#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged).  Usage: tools/gpu.sh <timeout-s> '<command>'
t=$1; shift
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 45
done
exit 3
