#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity (image, host layer, jpeg, api behaviour)"
timeout 900 python -m pytest tests/test_image_gpu.py tests/test_host_layer_gpu.py tests/test_jpeg_gpu.py tests/test_api_behaviour_gpu.py -x -q -m gpu -s 2>&1 | grep -E "passed|failed|jpeg ingest|Error" | tail -6
echo "== image timing"
timeout 300 python scripts/dev_image_bench.py 2>&1 | tail -4
echo "== bench (full, N=1)"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_n1.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', l['e2e']['value'], 'parity', l['parity_check']['ok'])
for k,v in l['secondary'].items(): print(k, json.dumps(v)[:700])
PY
