#!/bin/sh
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "busy" (exit 3)
T="$1"; shift
i=0
while [ $i -lt 30 ]; do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  i=$((i + 1)); sleep 90
done
exit 3
