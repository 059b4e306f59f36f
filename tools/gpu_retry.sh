#!/bin/bash
# usage: [GPUS=N] tools/gpu_retry.sh <timeout-seconds> '<command>'   -- retries gpurun while the pod has no free slot (exit 3)
T=$1; shift
G=""
if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
