#!/usr/bin/env bash
# tools/gpurun_retry.sh LOG TIMEOUT CMD — retry a gpurun call while the pod answers "transient" (no slot free; nothing charged).
LOG=$1; TO=$2; shift 2
for i in $(seq 1 40); do
  gpurun --timeout $TO -- "$@" > $LOG 2>&1
  grep -q "status=transient\|rc=3\|no box" $LOG || break
  sleep 120
done
echo "gpurun_retry: done after $i tries" >> $LOG
