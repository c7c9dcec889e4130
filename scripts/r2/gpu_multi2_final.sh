#!/bin/bash
# 2-GPU box, final tree: group scans through the boundary (single process, ncclCommInitAll; with and without bound exchanges)
# and the bench under torchrun (Hamming leg only), parity against the oracle over all 1 B rows
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_group_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -8
UCFP_GROUP_EXCHANGES=4 timeout 600 python -m pytest tests/test_group_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-images --no-paths --no-cpu-baseline > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
echo "bench rc=$?"
tail -3 gpurun_out/r2f_bench_n2.err
python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2f_bench_n2.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, 'e2e', l['e2e']['value'], 'parity', l.get('parity_check'))
    print('kernel_ms', l['roofline'].get('kernel_ms_per_step'))
except Exception as e: print('parse failed', e)
PY
