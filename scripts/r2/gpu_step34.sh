#!/bin/bash
# Round 2, step 34: two-group epilogue on the expansion form (3 producer warps, 20 warps): parity, 250M and 1B timings against lock-step
set -u
timeout 900 python -m pytest tests/test_hamming_gpu.py tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -3
UCFP_HAMMING_NO_OPS=1 timeout 600 python -m pytest tests/test_hamming_gpu.py -x -q -m gpu -k "tensor or config2" 2>&1 | tail -3
for E in 1 0; do
  echo "== STAGGER_EXP=$E"
  UCFP_HAMMING_STAGGER_EXP=$E timeout 300 python scripts/dev_hamming_bench.py 2.5e8 1024 2>&1 | tail -1
  UCFP_HAMMING_STAGGER_EXP=$E UCFP_HAMMING_NO_OPS=1 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 384,512,768,1024 2>&1 | tail -4
  UCFP_HAMMING_STAGGER_EXP=$E timeout 300 python scripts/dev_hamming_bench.py 1e9 1024 2>&1 | tail -1
done
