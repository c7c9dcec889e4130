#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (image, hamming, sharded, group)"
timeout 900 python -m pytest tests/test_image_gpu.py tests/test_hamming_gpu.py tests/test_sharded_gpu.py tests/test_group_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== image timing"
timeout 300 python scripts/dev_image_bench.py 2>&1 | tail -4
echo "== hamming timing"
timeout 200 python scripts/dev_hamming_bench.py 1.25e8 128,1024 2>&1 | tail -2
timeout 200 python scripts/dev_hamming_bench.py 1e9 1024 2>&1 | tail -1
echo "== ncu image 256"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:image_stream -c 1 -o gpurun_out/image_stream_256_v4 python scripts/prof_image.py 256 256 9472 > gpurun_out/ncu_img.log 2>&1; echo "ncu rc=$?"
