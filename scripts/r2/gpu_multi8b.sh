#!/bin/bash
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
echo "== group tests (single process, all $N GPUs)"
timeout 900 python -m pytest tests/test_group_gpu.py -x -q -m gpu 2>&1 | tail -3
run() {  # name, extra env
  local name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-paths --no-images --no-cpu-baseline --parity-queries 8 > gpurun_out/r2_bench_n${N}_$name.json 2> gpurun_out/r2_bench_n${N}_$name.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_bench_n${N}_$name.json').read().strip().splitlines()[-1])
    print('$name', {k:l[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', l['e2e']['value'], 'parity', l.get('parity_check',{}).get('ok'), 'kernel_ms', l['roofline'].get('kernel_ms_per_step'))
except Exception as e: print('parse failed', e)
PY
}
run ex0 UCFP_GROUP_EXCHANGES=0
run ex1 UCFP_GROUP_EXCHANGES=1
run ex2 UCFP_GROUP_EXCHANGES=2
run ex4 UCFP_GROUP_EXCHANGES=4
echo "== full bench N=$N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, 'e2e', l['e2e']['value'], 'parity', l.get('parity_check'))
    print('kernel_ms', l['roofline'].get('kernel_ms_per_step'), 'secondary', {k:(v.get('value'), v.get('result_crc'), v.get('parity_check',{}).get('ok')) for k,v in l['secondary'].items()})
except Exception as e: print('parse failed', e)
PY
