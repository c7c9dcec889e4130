#!/bin/bash
# Round-2 evidence run (one B200): full GPU test suite, contract bench line, robustness cases, ncu launch list and one full
# ncu capture per dominant kernel.  Every ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
echo "== full gpu suite"
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench N=1"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_n1.err
echo "== bench reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
echo "== robustness"
timeout 1500 python scripts/bench_robustness.py > gpurun_out/r2_robustness.jsonl 2> gpurun_out/r2_robustness.err; echo "robustness rc=$?"
cat gpurun_out/r2_robustness.jsonl | cut -c1-400
tail -3 gpurun_out/r2_robustness.err
echo "== launch list (bench at 2.5e8 rows)"
CMD="python bench.py --steps 1 --warmup 3 --codes 2.5e8 --no-cpu-baseline --no-images --no-paths --parity-queries 2"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_hamming.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
echo "== ncu full: hamming tensor scan"
CMD="python scripts/dev_hamming_bench.py 2.5e8 1024"
$CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_mma_scan_kernel -s 9 -c 1 -o gpurun_out/r2_hamming_mma_q1024 $CMD > gpurun_out/ncu_full.log 2>&1
echo "hamming q1024 rc=$?"
echo "== ncu full: jaccard"
CMD="python scripts/prof_scan.py jaccard 4e6 256"
$CMD > gpurun_out/plain5.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jaccard_scan -s 11 -c 1 -o gpurun_out/r2_jaccard_scan_q256 $CMD > gpurun_out/ncu_full4.log 2>&1
echo "jaccard rc=$?"
echo "== ncu full: cosine"
CMD="python scripts/prof_scan.py cosine 2e6 1024"
$CMD > gpurun_out/plain4.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cosine_coarse -s 11 -c 1 -o gpurun_out/r2_cosine_coarse_q1024 $CMD > gpurun_out/ncu_full3.log 2>&1
echo "cosine rc=$?"
echo "== ncu full: image 1024 / 256"
CMD="python scripts/prof_image.py 1024 1024 1184"
$CMD > gpurun_out/plain6.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:image_stream -s 2 -c 1 -o gpurun_out/r2_image_stream_1024 $CMD > gpurun_out/ncu_full5.log 2>&1
echo "image 1024 rc=$?"
CMD="python scripts/prof_image.py 256 256 9472"
$CMD > gpurun_out/plain7.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:image_stream -s 2 -c 1 -o gpurun_out/r2_image_stream_256 $CMD > gpurun_out/ncu_full6.log 2>&1
echo "image 256 rc=$?"
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
