#!/bin/bash
# Development aid: gpurun with retries on "no slot free" (exit 3).  usage: gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  gpurun "$@" > $log 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 120
done
exit 3
