#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <max_tries> <gpurun args...>   -- retries while the pod is busy (exit code 3)
log=$1; tries=$2; shift 2
for i in $(seq 1 $tries); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
