#!/bin/bash
# usage: tools/gpu_retry.sh <tag> <timeout_s> <command...>   -- retries gpurun while the pod answers "busy" (exit 3)
tag=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "attempt $attempt rc=$rc" >> gpurun_out/${tag}_call.log; exit $rc; fi
  sleep 45
done
exit 3
