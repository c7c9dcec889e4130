#!/bin/bash
# Round 2, step 36: which form serves 640-896 queries (stage images with the two-group epilogue vs expansion with it)
set -u
for Q in 640 768 896; do
  echo "== $Q queries: images / expansion"
  timeout 300 python scripts/dev_hamming_bench.py 2.5e8 $Q 2>&1 | tail -1
  UCFP_HAMMING_IMG_MAXQ=512 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 $Q 2>&1 | tail -1
done
timeout 300 python -m pytest tests/test_hamming_gpu.py -x -q -m gpu 2>&1 | tail -3
