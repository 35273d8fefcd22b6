#!/bin/bash
# usage: [GPUS=2] gpurun_retry.sh <timeout-seconds> <command...>   (retries while the pod answers "busy"; exit code 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' /root/repo/gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  sleep 90
done
exit 3
