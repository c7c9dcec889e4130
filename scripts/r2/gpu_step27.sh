#!/bin/bash
# Round 2, step 27: where the tensor scan's item time goes -- built with -DUCFP_HAMMING_DIAG: skip the min/max test (1) / also the TMEM loads (2)
set -u
for S in 0 1; do for D in 0 1 2; do
  echo "== STAGGER=$S DIAG=$D"
  UCFP_HAMMING_STAGGER=$S UCFP_HAMMING_DIAG=$D timeout 300 python scripts/dev_hamming_bench.py 2.5e8 1024 2>&1 | tail -1
  UCFP_HAMMING_STAGGER=$S UCFP_HAMMING_DIAG=$D UCFP_HAMMING_IMG_MAXQ=1024 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 1024 2>&1 | tail -1
done; done
