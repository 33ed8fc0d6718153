#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> <command...>   — retries while the pod answers busy (rc 3)
LOG=$1; TMO=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
