#!/bin/bash
# Round 2, step 30: two-group epilogue; second strip in flight while the first is reduced, release after the second has landed
set -u
echo "== parity, stagger + images at every batch size"
UCFP_HAMMING_STAGGER=1 UCFP_HAMMING_IMG_MAXQ=1024 timeout 600 python -m pytest tests/test_hamming_gpu.py tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -3
for D in 0 1 2; do
  echo "== STAGGER=1 images DIAG=$D"
  UCFP_HAMMING_STAGGER=1 UCFP_HAMMING_DIAG=$D UCFP_HAMMING_IMG_MAXQ=1024 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 128,256,512,1024 2>&1 | tail -4
done
echo "== STAGGER=0 product"
UCFP_HAMMING_STAGGER=0 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 128,256,512,1024 2>&1 | tail -4
