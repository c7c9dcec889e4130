#!/bin/bash
# Round 2, step 31: two-group epilogue as the default: whole GPU suite, dev timings, the contract bench on 1 GPU
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/dev_hamming_bench.py 2.5e8 64,128,256,512,1024 2>&1 | tail -5
timeout 300 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_step31.json 2> gpurun_out/bench_step31.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_step31.json
