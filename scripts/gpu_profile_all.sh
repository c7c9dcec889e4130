#!/bin/bash
# Runs on the GPU box (under gpurun): contract bench line, ncu launch list, one full ncu capture per dominant kernel.
# Every ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --codes 2.5e8 --no-cpu-baseline --no-images"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_hamming.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_scan_kernel -s 5 -c 1 -o gpurun_out/hamming_scan_q1024 $CMD > gpurun_out/ncu_full.log 2>&1
echo "hamming q1024 rc=$?"
CMD="python scripts/dev_hamming_bench.py 1e9 1"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hamming_scan_kernel -s 2 -c 1 -o gpurun_out/hamming_scan_q1 $CMD > gpurun_out/ncu_full2.log 2>&1
echo "hamming q1 rc=$?"
CMD="python scripts/prof_scan.py cosine 2e6 1024"
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cosine_coarse -s 11 -c 1 -o gpurun_out/cosine_coarse_q1024 $CMD > gpurun_out/ncu_full3.log 2>&1
echo "cosine rc=$?"
CMD="python scripts/prof_scan.py jaccard 4e6 256"
$CMD > gpurun_out/plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:jaccard_scan -s 11 -c 1 -o gpurun_out/jaccard_scan_q256 $CMD > gpurun_out/ncu_full4.log 2>&1
echo "jaccard rc=$?"
CMD="python scripts/prof_image.py 1024 1024 1184"
$CMD > gpurun_out/plain6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:image_stream -s 2 -c 1 -o gpurun_out/image_stream_1024 $CMD > gpurun_out/ncu_full5.log 2>&1
echo "image rc=$?"
python scripts/bench_paths.py > gpurun_out/paths.jsonl 2> gpurun_out/paths.err; echo "paths rc=$?"
