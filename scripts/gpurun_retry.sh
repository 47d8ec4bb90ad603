#!/bin/bash
# usage: gpurun_retry.sh <log> <gpurun args...>: retries while the pod answers busy (exit code 3)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then break; fi
  sleep 90
done
echo "gpurun_retry done rc=$rc try=$i" >> "$log"
