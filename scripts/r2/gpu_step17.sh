#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (image, host layer, jpeg)"
timeout 900 python -m pytest tests/test_image_gpu.py tests/test_host_layer_gpu.py tests/test_jpeg_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== image timing"
timeout 300 python scripts/dev_image_bench.py 2>&1 | tail -4
echo "== ncu image 256 / 1024"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:image_stream -c 1 -o gpurun_out/image_stream_256_v7 python scripts/prof_image.py 256 256 9472 > gpurun_out/ncu_img.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:image_stream -c 1 -o gpurun_out/image_stream_1024_v7 python scripts/prof_image.py 1024 1024 1184 > gpurun_out/ncu_img2.log 2>&1; echo "ncu rc=$?"
