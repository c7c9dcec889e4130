#!/bin/bash
# Runs on a GPU box with >= 2 GPUs (gpurun --gpus N): the multi-rank bench under torchrun at every N <= the GPUs present,
# each under its own timeout so that a hung collective cannot eat the budget.  First thing to run in a new round: the
# top-up loop of bench.py's clock sampler once ran a collective on some ranks only and deadlocked a 4-rank run.
set -u
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
for N in 2 4 8; do
    [ "$N" -le "$G" ] || continue
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29600 + N)) \
        bench.py --gpus "$N" --steps 5 --warmup 3 --no-images > "gpurun_out/bench_n$N.json" 2> "gpurun_out/bench_n$N.err"
    echo "N=$N rc=$? $(cut -c1-200 "gpurun_out/bench_n$N.json")"
done
