#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/r2/debug_multihash.py 2>&1 | tail -12
echo "== tests"
timeout 1200 python -m pytest tests/test_multihash_gpu.py tests/test_host_layer_gpu.py tests/test_jpeg_gpu.py tests/test_mutation_gpu.py -q -m gpu -s 2>&1 | tail -30
echo "== early poll timing (WAIT=4) vs default"
for w in 0 4; do UCFP_HAMMING_WAIT=$w timeout 200 python scripts/dev_hamming_bench.py 2.5e8 256,1024 2>&1 | tail -2; done
