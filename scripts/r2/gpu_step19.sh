#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== jaccard 6.25M x 256 (one of 8 shards)"
timeout 300 python scripts/dev_jaccard_timing.py 6.25e6 256 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_jaccard_shard.csv python scripts/dev_jaccard_timing.py 6.25e6 256 > gpurun_out/ncu_j.log 2>&1; echo "ncu rc=$?"
echo "== cosine 2.5M x 512 x 1024 queries (one of 8 shards)"
timeout 300 python scripts/dev_cosine_bench.py 2.5e6 1024 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_cosine_shard.csv python scripts/prof_scan.py cosine 2.5e6 1024 > gpurun_out/ncu_c.log 2>&1; echo "ncu rc=$?"
