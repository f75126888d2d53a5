#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   - retries while the pod answers "busy" (exit code 3)
LOG="$1"; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> "$LOG"; exit $rc; fi
  sleep 45
done
echo "rc=3 (gave up)" >> "$LOG"
