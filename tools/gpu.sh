#!/bin/bash
# usage: tools/gpu.sh [--gpus N] TIMEOUT 'command'   — gpurun with retries while the pod is busy (exit code 3)
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
