#!/bin/bash
# Round-2 final evidence run (one B200) of the shipped tree: full GPU suite, smoke, contract bench line + reference arm, robustness and
# flood cases, launch list and full ncu captures of the two forms of the tensor scan.  Every ncu run follows a plain run that exited 0.
set -u
mkdir -p gpurun_out
echo "== full gpu suite"
timeout 2400 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | tail -12
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench N=1"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"
tail -2 gpurun_out/r2f_bench_n1.err
echo "== bench reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"
echo "== robustness"
timeout 1500 python scripts/bench_robustness.py > gpurun_out/r2f_robustness.jsonl 2> gpurun_out/r2f_robustness.err; echo "robustness rc=$?"
cut -c1-300 gpurun_out/r2f_robustness.jsonl
echo "== flood"
timeout 300 python scripts/dev_flood_bench.py 5e7 1e6 1024 8 2>&1 | tail -1 | tee gpurun_out/r2f_flood.jsonl
UCFP_RESCAN_ROUNDS=0 timeout 300 python scripts/dev_flood_bench.py 5e7 1e6 1024 8 2>&1 | tail -1 | tee -a gpurun_out/r2f_flood.jsonl
timeout 300 python scripts/dev_flood_bench.py 5e7 1e6 1024 0 2>&1 | tail -1 | tee -a gpurun_out/r2f_flood.jsonl
echo "== dev sweep"
timeout 300 python scripts/dev_hamming_bench.py 2.5e8 1,2,16,64,128,256,512,640,768,1024 2>&1 | tee gpurun_out/r2f_hamming_sweep.jsonl | tail -10
timeout 300 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1 | tee -a gpurun_out/r2f_hamming_sweep.jsonl
echo "== launch list (bench at 2.5e8 rows)"
CMD="python bench.py --steps 1 --warmup 3 --codes 2.5e8 --no-cpu-baseline --no-images --no-paths --parity-queries 2"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2f_launches_hamming.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
echo "== ncu full: tensor scan, 1024 queries (expansion form, two-group epilogue)"
CMD="python scripts/dev_hamming_bench.py 2.5e8 1024"
$CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_mma_scan_kernel -s 9 -c 1 -o gpurun_out/r2f_hamming_mma_q1024 $CMD > gpurun_out/ncu_full.log 2>&1
echo "hamming q1024 rc=$?"
echo "== ncu full: tensor scan, 512 queries (image form, two-group epilogue)"
CMD="python scripts/dev_hamming_bench.py 2.5e8 512"
$CMD > gpurun_out/plain3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hamming_mma_scan_kernel -s 9 -c 1 -o gpurun_out/r2f_hamming_mma_q512 $CMD > gpurun_out/ncu_full2.log 2>&1
echo "hamming q512 rc=$?"
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
