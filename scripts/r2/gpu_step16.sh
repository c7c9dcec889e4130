#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (image, host layer, jpeg)"
timeout 900 python -m pytest tests/test_image_gpu.py tests/test_host_layer_gpu.py tests/test_jpeg_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== image timing: default build"
timeout 300 python scripts/dev_image_bench.py 2>&1 | tail -4
echo "== compute-sanitizer memcheck on the small cases"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/sanitize_memcheck.log
echo "== ncu image 256 (default build)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:image_stream -c 1 -o gpurun_out/image_stream_256_v6 python scripts/prof_image.py 256 256 9472 > gpurun_out/ncu_img.log 2>&1; echo "ncu rc=$?"
echo "== image timing: single row loop (no run fast path)"
UCFP_BUILD_DEFINES="-DUCFP_IMG_SINGLE_LOOP" python -m ucfp_b200.build --force > /dev/null 2>&1; echo "build rc=$?"
timeout 300 python -m pytest tests/test_image_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/dev_image_bench.py 2>&1 | tail -4
