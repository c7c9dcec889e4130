#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list, full ncu captures of the two Hamming scan kernels.
set -u
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_hamming.json 2> gpurun_out/bench_hamming.err
echo "bench rc=$?"; tail -c 4000 gpurun_out/bench_hamming.json
CMD="python bench.py --steps 1 --warmup 3 --codes 2.5e8 --no-cpu-baseline --no-images"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_hamming.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_mma_scan_kernel -s 19 -c 1 -o gpurun_out/hamming_mma_q1024 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
# the stage-image variant of the tensor scan (batches of 64-896 queries): 256 queries over 250 M codes, largest launch
CMD2="python bench.py --steps 1 --warmup 3 --codes 2.5e8 --queries 256 --no-cpu-baseline --no-images"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_mma_scan_kernel -s 19 -c 1 -o gpurun_out/hamming_mma_q256 $CMD2 > gpurun_out/ncu_full2.log 2>&1
echo "ncu full (q256) rc=$?"; tail -2 gpurun_out/ncu_full2.log
