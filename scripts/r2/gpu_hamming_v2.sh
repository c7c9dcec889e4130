#!/bin/bash
# Round 2, step 1: parity of the second-generation tensor scan and a timing sweep of its variants.
set -u
mkdir -p gpurun_out
for W in 16 8; do
  echo "== parity, EPI_WARPS=$W"
  UCFP_HAMMING_EPI_WARPS=$W timeout 600 python -m pytest tests/test_hamming_gpu.py -x -q -m gpu 2>&1 | tail -3
done
for cfg in "1 16" "2 16" "2 8"; do
  set -- $cfg
  echo "== timing MMA_V=$1 EPI_WARPS=$2"
  UCFP_HAMMING_MMA_V=$1 UCFP_HAMMING_EPI_WARPS=$2 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 64,128,256,512,1024 2>&1 | tail -5
done
cd scripts/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && timeout 60 ./pipe_rates | tee ../../gpurun_out/pipe_rates.txt
