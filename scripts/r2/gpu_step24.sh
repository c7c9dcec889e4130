#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (cosine, group, host layer, api, config4)"
timeout 1200 python -m pytest tests/test_cosine_gpu.py tests/test_group_gpu.py tests/test_host_layer_gpu.py tests/test_api_behaviour_gpu.py tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_configs_gpu.py -x -q -m gpu -k "config4" 2>&1 | tail -3
echo "== cosine shard timing (2.5M x 512): probe / seed"
timeout 300 python scripts/dev_cosine_bench.py 2.5e6 2>&1 | tail -3
UCFP_COSINE_NO_PROBE=1 timeout 300 python scripts/dev_cosine_bench.py 2.5e6 2>&1 | tail -1
echo "== cosine 20M timing: probe / seed"
timeout 600 python scripts/dev_cosine_bench.py 2e7 2>&1 | tail -2
UCFP_COSINE_NO_PROBE=1 timeout 600 python scripts/dev_cosine_bench.py 2e7 2>&1 | tail -1
echo "== launch list cosine shard"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_cosine_shard3.csv python scripts/prof_scan.py cosine 2.5e6 1024 > gpurun_out/ncu_c.log 2>&1; echo "ncu rc=$?"
