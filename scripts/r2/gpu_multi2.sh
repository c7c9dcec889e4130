#!/bin/bash
# 2-GPU box: group scans through the boundary (single process, ncclCommInitAll) and the bench under torchrun (ncclCommInitRank)
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== group tests (single process, all GPUs of the box)"
timeout 900 python -m pytest tests/test_group_gpu.py -x -q -m gpu 2>&1 | tail -15
echo "== bench N=2 under torchrun"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, 'e2e', l['e2e']['value'], 'parity', l.get('parity_check'))
    print('kernel_ms', l['roofline'].get('kernel_ms_per_step'), 'secondary', {k:(v.get('value'), v.get('result_crc'), v.get('parity_check',{}).get('ok')) for k,v in l['secondary'].items()})
except Exception as e: print('parse failed', e)
PY
