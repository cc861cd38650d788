#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> <command...> : retry while the pod answers "transient" (nothing charged)
LOG=$1; TO=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if grep -q "status=transient" $LOG; then sleep 90; else break; fi
done
