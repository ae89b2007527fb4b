#!/bin/bash
# usage: tools/gpu_retry.sh LOGFILE [--gpus N] --timeout S -- 'command'   (retries while the pod answers busy)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 45
done
exit 3
