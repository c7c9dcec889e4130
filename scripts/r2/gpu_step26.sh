#!/bin/bash
# Round 2, step 26: the tensor scan's epilogue in two groups out of phase (UCFP_HAMMING_STAGGER=1): parity, then timing against the product schedule.
set -u
mkdir -p gpurun_out
echo "== parity, stagger"
UCFP_HAMMING_STAGGER=1 timeout 600 python -m pytest tests/test_hamming_gpu.py tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -3
for S in 0 1; do
  echo "== timing STAGGER=$S (expansion at 1024, images below)"
  UCFP_HAMMING_STAGGER=$S timeout 300 python scripts/dev_hamming_bench.py 2.5e8 64,128,256,512,1024 2>&1 | tail -5
  echo "== timing STAGGER=$S, images at 1024"
  UCFP_HAMMING_STAGGER=$S UCFP_HAMMING_IMG_MAXQ=1024 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 1024 2>&1 | tail -1
done
