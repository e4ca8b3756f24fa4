#!/bin/bash
# gpurun wrapper: retries while the pod answers "transient"/busy (nothing charged). usage: scripts/gpu.sh <timeout_s> '<command>'
T=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|exit code 3"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
