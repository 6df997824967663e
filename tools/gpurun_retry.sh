#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / status=transient).  Usage: tools/gpurun_retry.sh LOG [gpurun args...] -- 'command'
LOG=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if ! grep -q "status=transient\|status=busy" "$LOG" && [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
