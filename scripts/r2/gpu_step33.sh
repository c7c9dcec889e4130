#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_hamming_gpu.py -x -q -m gpu -k "large_k_takes" 2>&1 | tail -40
for i in 1 2; do
UCFP_RESCAN_ROUNDS=0 timeout 300 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1
timeout 300 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1
done
