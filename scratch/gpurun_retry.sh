#!/bin/bash
# usage: scratch/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "busy / transient" (nothing charged)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > /tmp/gpurun_retry.out 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "retry in a few minutes" /tmp/gpurun_retry.out && ! python -c "import json,sys; d=json.load(open('/root/repo/gpurun_out/.last_call.json')); sys.exit(0 if d.get('status')=='transient' else 1)" 2>/dev/null; then
    cat /tmp/gpurun_retry.out; exit $rc
  fi
  sleep 75
done
cat /tmp/gpurun_retry.out; exit 3
