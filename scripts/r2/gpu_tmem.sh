#!/bin/bash
set -u
mkdir -p gpurun_out
cd scripts/micro && timeout 120 ./tmem_port | tee ../../gpurun_out/tmem_port.txt; cd ../..
for W in 16 8; do
  echo "== parity, EPI_WARPS=$W"
  UCFP_HAMMING_EPI_WARPS=$W timeout 600 python -m pytest tests/test_hamming_gpu.py -x -q -m gpu 2>&1 | tail -3
done
