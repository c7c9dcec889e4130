#!/bin/bash
# Round 2, step 3: whole GPU suite after the lane / mutation / batcher / group refactor, then wait-mode timing of the tensor scans.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | tail -25
for cfg in "1 0" "1 1" "1 3" "2 0" "2 1" "2 3"; do
  set -- $cfg
  echo "== timing MMA_V=$1 WAIT=$2"
  UCFP_HAMMING_MMA_V=$1 UCFP_HAMMING_WAIT=$2 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 256,1024 2>&1 | tail -2
done
