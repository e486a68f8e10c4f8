#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE TIMEOUT [--gpus N] -- 'command'   (retries while the pod answers busy: rc 3 / transient)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1; rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
exit $rc
