#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== MMA_V=3 EPI_WARPS=16 (setmaxnreg) parity:"; UCFP_HAMMING_MMA_V=3 timeout 300 python -m pytest tests/test_hamming_gpu.py -x -q -m gpu -k "tensor or config2" 2>&1 | tail -2
echo "   timing"
UCFP_HAMMING_MMA_V=3 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 128,256,1024 2>&1 | tail -3
echo "== multihash / host layer / jpeg"
timeout 900 python -m pytest tests/test_multihash_gpu.py tests/test_host_layer_gpu.py tests/test_jpeg_gpu.py -q -m gpu -s 2>&1 | tail -60
