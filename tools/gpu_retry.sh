#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'   (retries while the pod answers "busy")
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.txt; then sleep 90; continue; fi
  break
done
tail -12 /tmp/gpurun_last.txt
exit $rc
