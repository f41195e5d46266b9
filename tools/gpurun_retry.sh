#!/bin/bash
# gpurun with retries while the pod answers "transient / busy" (nothing is charged for those).  Usage: gpurun_retry.sh [gpurun args]
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  echo "$out"
  if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  exit $rc
done
exit 3
