#!/usr/bin/env bash
# usage: scripts/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers busy / transient)
T="$1"; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  echo "$out"
  if echo "$out" | grep -q "status=transient\|retry in a few minutes" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
