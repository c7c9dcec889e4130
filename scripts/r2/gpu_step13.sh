#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (hamming, sharded, group, multihash, mutation)"
timeout 900 python -m pytest tests/test_hamming_gpu.py tests/test_sharded_gpu.py tests/test_group_gpu.py tests/test_multihash_gpu.py tests/test_mutation_gpu.py -x -q -m gpu 2>&1 | tail -4
for sp in 1 0; do
  if [ $sp = 1 ]; then export UCFP_HAMMING_NO_SPILL=1; echo "== settle in place"; else unset UCFP_HAMMING_NO_SPILL; echo "== parked strips"; fi
  timeout 200 python scripts/dev_hamming_bench.py 1.25e8 128,1024 2>&1 | tail -2
  timeout 200 python scripts/dev_hamming_bench.py 1e9 1024 2>&1 | tail -1
done
echo "== growth 4 / 16 with parked strips"
UCFP_HAMMING_GROWTH=4 timeout 200 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1
UCFP_HAMMING_GROWTH=4 timeout 200 python scripts/dev_hamming_bench.py 1e9 1024 2>&1 | tail -1
UCFP_HAMMING_GROWTH=16 timeout 200 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1
echo "== launch list at 1.25e8 rows (ncu, serialised)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_125m_parked.csv python scripts/dev_hamming_bench.py 1.25e8 1024 > gpurun_out/ncu_list2.log 2>&1; echo "ncu rc=$?"
echo "== full gpu suite"
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
