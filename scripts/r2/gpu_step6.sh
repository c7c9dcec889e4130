#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "4 16" "4 8"; do
  set -- $cfg
  echo "== MMA_V=$1 EPI_WARPS=$2 parity:"; UCFP_HAMMING_MMA_V=$1 UCFP_HAMMING_EPI_WARPS=$2 timeout 600 python -m pytest tests/test_hamming_gpu.py -x -q -m gpu -k "tensor or config2" 2>&1 | tail -2
  echo "   timing"
  UCFP_HAMMING_MMA_V=$1 UCFP_HAMMING_EPI_WARPS=$2 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 128,256,1024 2>&1 | tail -3
done
CMD="python scripts/dev_hamming_bench.py 1.2e8 1024"
UCFP_HAMMING_MMA_V=4 UCFP_HAMMING_EPI_WARPS=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:hamming_mma_scan4 -s 4 -c 1 -o gpurun_out/r2_mma4_w8 -f $CMD > gpurun_out/ncu_v4.log 2>&1
echo "ncu rc=$?"
