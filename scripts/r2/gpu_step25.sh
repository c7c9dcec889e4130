#!/bin/bash
set -u
for cfg in "2048 8" "2048 16" "4096 8" "4096 16" "8192 8" "8192 4" "1024 8"; do
  set -- $cfg
  echo "== probe rows=$1 first chunk = rows x $2"
  UCFP_COSINE_PROBE_ROWS=$1 UCFP_COSINE_PROBE_GROWTH=$2 timeout 300 python scripts/dev_cosine_bench.py 2.5e6 2>&1 | tail -1
  UCFP_COSINE_PROBE_ROWS=$1 UCFP_COSINE_PROBE_GROWTH=$2 timeout 300 python scripts/dev_cosine_bench.py 2e7 2>&1 | tail -1
done
echo "== exhaustive seed"
UCFP_COSINE_NO_PROBE=1 timeout 300 python scripts/dev_cosine_bench.py 2.5e6 2>&1 | tail -1
UCFP_COSINE_NO_PROBE=1 timeout 300 python scripts/dev_cosine_bench.py 2e7 2>&1 | tail -1
