#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "1 16" "2 16" "2 8" "3 16"; do
  set -- $cfg
  echo "== timing MMA_V=$1 EPI_WARPS=$2"
  UCFP_HAMMING_MMA_V=$1 UCFP_HAMMING_EPI_WARPS=$2 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 128,256,1024 2>&1 | tail -3
  echo "   parity:"; UCFP_HAMMING_MMA_V=$1 UCFP_HAMMING_EPI_WARPS=$2 timeout 600 python -m pytest tests/test_hamming_gpu.py -x -q -m gpu -k "tensor or config2" 2>&1 | tail -1
done
