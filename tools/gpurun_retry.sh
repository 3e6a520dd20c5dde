#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged).  usage: gpurun_retry.sh <timeout> <out file> <command...>
t=$1; out=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
