#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== v3 timing"
UCFP_HAMMING_MMA_V=3 timeout 300 python scripts/dev_hamming_bench.py 2.5e8 256,1024 2>&1 | tail -2
echo "== v3 parity"
UCFP_HAMMING_MMA_V=3 timeout 600 python -m pytest tests/test_hamming_gpu.py -x -q -m gpu -k "tensor or config2" 2>&1 | tail -2
echo "== bench N=1"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; echo "rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_n1_a.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l.get('parity_check'), l['roofline']['frac'], l['roofline']['kernel_ms_per_step'])
PY
tail -5 gpurun_out/r2_bench_n1_a.err
