#!/bin/bash
# Round 2, step 32: re-scan rounds for overflowed candidate lists (Hamming, Jaccard) + sampling append
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hamming_gpu.py tests/test_jaccard_gpu.py tests/test_multihash_gpu.py tests/test_sharded_gpu.py tests/test_group_gpu.py tests/test_mutation_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== flood: 50M rows, 1M identical rows with descending ids, 8 of 1024 queries hit it"
timeout 300 python scripts/dev_flood_bench.py 5e7 1e6 1024 8 2>&1 | tail -1
UCFP_RESCAN_ROUNDS=0 timeout 300 python scripts/dev_flood_bench.py 5e7 1e6 1024 8 2>&1 | tail -1
echo "== no flood hit (fast path cost of the extra launches)"
timeout 300 python scripts/dev_flood_bench.py 5e7 1e6 1024 0 2>&1 | tail -1
UCFP_RESCAN_ROUNDS=0 timeout 300 python scripts/dev_flood_bench.py 5e7 1e6 1024 0 2>&1 | tail -1
timeout 300 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1
timeout 300 python scripts/dev_jaccard_timing.py 6.25e6 256 2>&1 | tail -1
