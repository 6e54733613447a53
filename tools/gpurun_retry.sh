#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  usage: gpurun_retry.sh <log> <gpurun args...>
log="$1"; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$log" 2>&1; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
