#!/bin/bash
# gpurun with retries while the pod answers "busy / transient" (nothing is charged for those).
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  rc=$?
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|retry in a few minutes" || [ $rc -eq 3 ]; then
    sleep 90
    continue
  fi
  exit $rc
done
exit 3
