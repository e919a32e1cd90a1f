#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3, nothing charged).  usage: scripts/gpurun_retry.sh LOG TIMEOUT 'command'
LOG=$1; TMO=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$TMO" -- "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "gpurun_retry: done rc=$rc" >> "$LOG"
