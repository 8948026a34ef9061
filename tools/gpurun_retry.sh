#!/bin/bash
# tools/gpurun_retry.sh TIMEOUT [--gpus N] 'command'  -  gpurun with retries while the pod answers "busy" (exit code 3)
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T $G -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
