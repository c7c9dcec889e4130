#!/bin/bash
# usage: gpurun_retry.sh <timeout> [--gpus N] <script>   -- retries while the pod answers "transient" (nothing charged)
T=$1; shift
for attempt in $(seq 1 20); do
  out=$(gpurun --timeout "$T" "$@" 2>&1)
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient\|status=busy"; then echo "[retry $attempt] pod busy, sleeping"; sleep 150; continue; fi
  break
done
