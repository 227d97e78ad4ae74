#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <gpurun args...>   - retries while the pod answers busy/transient
log=$1; shift
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" "$log"; then sleep 60; continue; fi
  break
done
