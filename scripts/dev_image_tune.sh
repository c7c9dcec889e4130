#!/bin/bash
# developer sweep of the image kernel's shared-memory plan
for cfg in "0 0 0" "14 6 2" "8 6 2" "14 3 2" "26 6 2" "26 12 2" "14 6 3" "40 24 3"; do
  set -- $cfg
  echo "rowbuf=$1 stage=$2 stages=$3"
  UCFP_IMG_ROWBUF_KB=$1 UCFP_IMG_STAGE_KB=$2 UCFP_IMG_STAGES=$3 timeout 100 python scripts/dev_image_bench.py 2>&1 | head -2 | cut -c1-90
done
