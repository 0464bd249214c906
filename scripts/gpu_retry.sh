#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free): scripts/gpu_retry.sh <timeout> '<command>'
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
