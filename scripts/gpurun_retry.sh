#!/bin/bash
# Retries a gpurun call while the pod answers "busy" (exit 3: nothing charged).  Usage:
#   scripts/gpurun_retry.sh [--gpus N] <timeout-seconds> '<command>'
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] busy (attempt $i), sleeping 90 s"
  sleep 90
done
exit 3
