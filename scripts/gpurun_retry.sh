#!/usr/bin/env bash
# usage: gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod answers "busy" (exit 3)
LOG=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [[ $rc != 3 ]] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 90
done
exit 3
