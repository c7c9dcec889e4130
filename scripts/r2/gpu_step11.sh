#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== 125M-row shard (what one of 8 ranks scans): tensor-scan entry point sweep"
for mr in 65536 524288 4194304 33554432; do
  echo "MMA_MIN_ROWS=$mr"; UCFP_HAMMING_MMA_MIN_ROWS=$mr timeout 200 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1
done
for g in 4 16; do
  echo "GROWTH=$g"; UCFP_HAMMING_GROWTH=$g timeout 200 python scripts/dev_hamming_bench.py 1.25e8 1024 2>&1 | tail -1
done
echo "== launch list at 1.25e8 rows (ncu, serialised)"
CMD="python scripts/dev_hamming_bench.py 1.25e8 1024"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_125m.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu rc=$?"
echo "== N=1 bench (full)"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"
tail -2 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', l['e2e']['value'], 'parity', l.get('parity_check',{}).get('ok'), 'frac', l['roofline']['frac'])
print({k:(v.get('value'), v.get('parity_check')) for k,v in l['secondary'].items()})
print(l.get('cpu_baseline'))
PY
