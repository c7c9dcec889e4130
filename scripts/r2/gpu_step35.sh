#!/bin/bash
# Round 2, step 35: min/max chains rebalanced to the 16-operation minimum per test; A/B against the previous build on one box
set -u
timeout 900 python -m pytest tests/test_hamming_gpu.py tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -15
for L in "" ucfp_b200/libucfp_cuda_prev.so "" ucfp_b200/libucfp_cuda_prev.so; do
  echo "== lib=${L:-new}"
  UCFP_CUDA_LIB=$L timeout 300 python scripts/dev_hamming_bench.py 2.5e8 128,512,1024 2>&1 | tail -3
done
UCFP_CUDA_LIB= timeout 300 python scripts/dev_hamming_bench.py 1e9 1024 2>&1 | tail -1
UCFP_CUDA_LIB=ucfp_b200/libucfp_cuda_prev.so timeout 300 python scripts/dev_hamming_bench.py 1e9 1024 2>&1 | tail -1
