#!/bin/bash
# usage: scratch/grun.sh <tag> <timeout_s> [--gpus N] -- '<command>'   -> gpurun_out/<tag>_call.log ; retries while the pod answers busy (rc 3)
tag=$1; to=$2; shift 2
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$1" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/${tag}_call.log; then break; fi
  sleep 45
done
echo "rc=$rc" >> gpurun_out/${tag}_call.log
