#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (image, hamming, jaccard, sharded, group, multihash)"
timeout 900 python -m pytest tests/test_image_gpu.py tests/test_hamming_gpu.py tests/test_jaccard_gpu.py tests/test_sharded_gpu.py tests/test_group_gpu.py tests/test_multihash_gpu.py tests/test_mutation_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== image timing: default"
timeout 300 python scripts/dev_image_bench.py 2>&1 | tail -4
echo "== image timing: one column per thread"
UCFP_IMG_NO_CPT2=1 timeout 300 python scripts/dev_image_bench.py 2>&1 | tail -4 | head -1
for cfg in "8 6 2" "8 3 2" "6 3 2" "14 3 2" "10 4 3"; do
  set -- $cfg
  echo "rowbuf=$1 stage=$2 stages=$3"
  UCFP_IMG_ROWBUF_KB=$1 UCFP_IMG_STAGE_KB=$2 UCFP_IMG_STAGES=$3 timeout 100 python scripts/dev_image_bench.py 2>&1 | tail -4 | head -1
done
echo "== hamming timing"
timeout 200 python scripts/dev_hamming_bench.py 1.25e8 128,1024 2>&1 | tail -2
timeout 200 python scripts/dev_hamming_bench.py 1e9 1024 2>&1 | tail -1
echo "== jaccard timing"
timeout 300 python scripts/dev_jaccard_bench.py 2>&1 | tail -3
echo "== ncu image 256"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:image_stream -c 1 -o gpurun_out/image_stream_256_v5 python scripts/prof_image.py 256 256 9472 > gpurun_out/ncu_img.log 2>&1; echo "ncu rc=$?"
