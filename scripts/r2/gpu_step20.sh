#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (jaccard, sharded, group, host layer, configs jaccard)"
timeout 1200 python -m pytest tests/test_jaccard_gpu.py tests/test_sharded_gpu.py tests/test_group_gpu.py tests/test_host_layer_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_configs_gpu.py -x -q -m gpu -k "config3" 2>&1 | tail -3
echo "== jaccard shard timing: probe vs no probe"
timeout 300 python scripts/dev_jaccard_timing.py 6.25e6 256 2>&1 | tail -1
UCFP_JACCARD_NO_PROBE=1 timeout 300 python scripts/dev_jaccard_timing.py 6.25e6 256 2>&1 | tail -1
echo "== jaccard 50M timing: probe vs no probe"
timeout 600 python scripts/dev_jaccard_timing.py 5e7 256 2>&1 | tail -1
UCFP_JACCARD_NO_PROBE=1 timeout 600 python scripts/dev_jaccard_timing.py 5e7 256 2>&1 | tail -1
echo "== robustness (jaccard duplicates etc.)"
timeout 1500 python scripts/bench_robustness.py > gpurun_out/r2_robustness.jsonl 2> gpurun_out/r2_robustness.err; echo "robustness rc=$?"
cut -c1-330 gpurun_out/r2_robustness.jsonl
echo "== launch list jaccard shard with probe"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_jaccard_shard_probe.csv python scripts/dev_jaccard_timing.py 6.25e6 256 > gpurun_out/ncu_j.log 2>&1; echo "ncu rc=$?"
