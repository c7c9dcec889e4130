#!/bin/bash
set -u
mkdir -p gpurun_out
# prof_scan cosine: 3 scans, each 4-5 coarse launches (4096, 32K, 262K, 2.1M, rest).  Capture launches 0,1,2 of the LAST scan (skip 10).
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cosine_coarse -s 10 -c 3 -o gpurun_out/cosine_coarse_small python scripts/prof_scan.py cosine 2.5e6 1024 > gpurun_out/ncu_cs.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cosine_rescore -s 22 -c 4 -o gpurun_out/cosine_rescore_small python scripts/prof_scan.py cosine 2.5e6 1024 > gpurun_out/ncu_cr.log 2>&1; echo "ncu rc=$?"
