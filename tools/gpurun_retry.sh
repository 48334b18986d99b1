#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>'   -- retries while the pod answers busy / transient
T=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|no box\|busy"; then sleep 90; continue; fi
  break
done
