#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "3 8" "3 16"; do
  set -- $cfg
  echo "== MMA_V=$1 EPI_WARPS=$2 parity:"; UCFP_HAMMING_MMA_V=$1 UCFP_HAMMING_EPI_WARPS=$2 timeout 600 python -m pytest tests/test_hamming_gpu.py -x -q -m gpu -k "tensor or config2" 2>&1 | tail -2
  echo "   timing"
  UCFP_HAMMING_MMA_V=$1 UCFP_HAMMING_EPI_WARPS=$2 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 128,256,1024 2>&1 | tail -3
done
echo "== new tests"
timeout 900 python -m pytest tests/test_multihash_gpu.py -x -q -m gpu 2>&1 | tail -5
