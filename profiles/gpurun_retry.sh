#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged).  usage: gpurun_retry.sh <timeout_s> '<command>' [extra gpurun args]
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
