#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 / "transient"): tools/gpurun_retry.sh <log> <timeout> [--gpus N] -- <command>
log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then exit 0; fi
  sleep 90
done
exit 3
