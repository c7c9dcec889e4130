#!/bin/bash
# Round 2, step 41: flagged queries leave the tensor scan's hot test; three re-scan rounds
set -u
timeout 900 python -m pytest tests/test_hamming_gpu.py tests/test_jaccard_gpu.py tests/test_multihash_gpu.py tests/test_sharded_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -8
timeout 300 python scripts/dev_flood_bench.py 5e7 1e6 1024 8 2>&1 | tail -1
timeout 300 python scripts/dev_flood_bench.py 5e7 1e6 1024 64 2>&1 | tail -1
timeout 300 python scripts/dev_flood_bench.py 5e7 1e6 1024 0 2>&1 | tail -1
timeout 300 python scripts/r2/debug_flood.py 2>&1 | awk '{print $2, $3, $4, $5, $6, $7, $NF}' | sort | uniq -c | sort -rn | head -5
UCFP_RESCAN_ROUNDS=0 timeout 300 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1
timeout 300 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1
