#!/bin/bash
# Round 2, step 38: low-field view by rotation (SHF.L.W, ALU pipe) instead of IMAD.SHL (FMA pipe); A/B/C on one box
set -u
timeout 900 python -m pytest tests/test_hamming_gpu.py tests/test_sharded_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -15
for L in "" ucfp_b200/libucfp_cuda_prev.so ucfp_b200/libucfp_cuda_m2.so "" ucfp_b200/libucfp_cuda_prev.so ucfp_b200/libucfp_cuda_m2.so; do
  echo "== lib=${L:-new (rotate)}"
  UCFP_CUDA_LIB=$L timeout 300 python scripts/dev_hamming_bench.py 2.5e8 128,512,1024 2>&1 | tail -3
done
for L in "" ucfp_b200/libucfp_cuda_prev.so ucfp_b200/libucfp_cuda_m2.so; do
  UCFP_CUDA_LIB=$L timeout 300 python scripts/dev_hamming_bench.py 1e9 1024 2>&1 | tail -1
done
