#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (cosine, group, host layer, config4)"
timeout 1200 python -m pytest tests/test_cosine_gpu.py tests/test_group_gpu.py tests/test_host_layer_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_configs_gpu.py -x -q -m gpu -k "config4" 2>&1 | tail -3
echo "== cosine shard timing (2.5M x 512)"
timeout 300 python scripts/dev_cosine_bench.py 2.5e6 2>&1 | tail -3
echo "== cosine 20M timing"
timeout 600 python scripts/dev_cosine_bench.py 2e7 2>&1 | tail -2
echo "== launch list cosine shard"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_cosine_shard2.csv python scripts/prof_scan.py cosine 2.5e6 1024 > gpurun_out/ncu_c.log 2>&1; echo "ncu rc=$?"
