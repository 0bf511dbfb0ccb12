#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE TIMEOUT -- command...   (retries while the pod answers "transient / busy")
LOG=$1; TMO=$2; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|rc=3\b" $LOG || grep -q "retry in a few minutes" $LOG; then sleep 150; continue; fi
  break
done
