#!/bin/bash
# tools/gpurun_retry.sh LOG TIMEOUT CMD...: retries a gpurun call while the pod answers busy/transient
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|status=busy\|exit code 3\|no box" $LOG; then sleep 120; continue; fi
  break
done
