#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <log-file> '<command>'   -- retries while the pod has no free GPU slot (rc 3)
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
