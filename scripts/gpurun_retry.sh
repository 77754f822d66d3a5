#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> '<command>' — gpurun with retries while the pod answers "transient" (no box free)
T=$1; shift
for attempt in $(seq 1 20); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; exit 3
