#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python scripts/dev_hamming_bench.py 1.2e8 1024"
UCFP_HAMMING_MMA_V=2 $CMD > gpurun_out/plain_v2.log 2>&1 || exit 1
UCFP_HAMMING_MMA_V=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:hamming_mma_scan2 -s 4 -c 1 -o gpurun_out/r2_mma2_w16 -f $CMD > gpurun_out/ncu_v2.log 2>&1
echo "ncu v2 rc=$?"; tail -2 gpurun_out/ncu_v2.log
UCFP_HAMMING_MMA_V=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:hamming_mma_scan_kernel -s 4 -c 1 -o gpurun_out/r2_mma1 -f $CMD > gpurun_out/ncu_v1.log 2>&1
echo "ncu v1 rc=$?"; tail -2 gpurun_out/ncu_v1.log
