#!/bin/bash
# gpurun with retries while the pod answers "transient" (no box / slot free; nothing charged).
# usage: tools/gpurun_retry.sh [--gpus N] TIMEOUT 'command'
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 100; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
