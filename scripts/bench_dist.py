"""Multi-GPU runs of the non-headline configs (BASELINE.json configs 3, 4, 5), one process per GPU under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_dist.py jaccard|cosine|image

Same protocol as bench.py: corpus sharded by record range (strong scaling: the config's total size at every N),
every rank scans its slice, NCCL all-gather of the per-rank top-k, identical merge on every rank; images split
across ranks with no exchange.  CUDA-event timing, barrier + synchronize on both sides, max over ranks."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ucfp_b200 import Context, Corpus, _ffi  # noqa: E402
from ucfp_b200.sharding import shard_range  # noqa: E402

what = sys.argv[1]
small = "--small" in sys.argv
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = Context(local)
K, STEPS = 10, 5


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps):
    for _ in range(3):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / steps


def splitmix(seed, idx):
    z = np.uint64(seed) ^ ((idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def view(corpus, shape, typestr):
    class _A:
        __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (corpus.device_rows_ptr(), False), "version": 2}
    return torch.as_tensor(_A(), device=dev)


line = None
with np.errstate(over="ignore"):
    if what == "jaccard":
        n_total, nq = (5_000_000 if small else 50_000_000), 256
        lo, hi = shard_range(n_total, rank, world)
        n = hi - lo
        corpus = Corpus(ctx, _ffi.KIND_MINHASH128, n)
        corpus.set_id_base(lo)
        corpus.append_synthetic(0x5EED, lo, n)
        q = splitmix(77, np.arange(nq * 128)).reshape(nq, 128)
        rng = np.random.default_rng(0)                   # same stream on every rank; each plants the rows it owns
        v = view(corpus, (n, 128), "<i8")
        for c0 in range(0, n_total // 100, 100_000):
            m = min(100_000, n_total // 100 - c0)
            rows = rng.choice(n_total, m, replace=False)
            qi, p = rng.integers(0, nq, m), rng.choice([0.9, 0.7, 0.5], m)
            mask = rng.random((m, 128)) < p[:, None]
            mine = (rows >= lo) & (rows < hi)
            if mine.any():
                base = splitmix(99 + c0, np.arange(m * 128)).reshape(m, 128)
                base[mask] = q[qi][mask]
                v[torch.from_numpy(rows[mine] - lo).to(dev)] = torch.from_numpy(base[mine].view(np.int64)).to(dev)
        corpus.refresh()
        qd = torch.from_numpy(q.view(np.int64)).to(dev)
        il, kl = torch.empty((nq, K), dtype=torch.int64, device=dev), torch.empty((nq, K), dtype=torch.int32, device=dev)
        ia, ka = torch.empty((world, nq, K), dtype=torch.int64, device=dev), torch.empty((world, nq, K), dtype=torch.int32, device=dev)
        io, ko = torch.empty_like(il), torch.empty_like(kl)

        def step():
            if world == 1:
                corpus.scan_jaccard(qd, K, io, ko)
            else:
                corpus.scan_jaccard(qd, K, il, kl)
                dist.all_gather_into_tensor(ia, il)
                dist.all_gather_into_tensor(ka, kl)
                ctx.merge_topk_u32(ia, ka, world, nq, K, True, io, ko)
        ms = timed(step, STEPS)
        assert int(ko[:, 0].min()) >= 100, "planted near-duplicates not found"
        line = {"path": "jaccard", "metric": "queries/s", "value": nq / ms * 1e3, "ms_per_step": ms, "n_gpus": world, "scaling": "strong",
                "config": {"workload": f"MinHash-128 Jaccard top-{K}, {nq}-query batch over {n_total} signatures (1 % planted)", "rows_per_gpu": n}}
    elif what == "cosine":
        n_total, dim, nq = (2_000_000 if small else 20_000_000), 512, 1024
        lo, hi = shard_range(n_total, rank, world)
        n = hi - lo
        corpus = Corpus(ctx, _ffi.KIND_COSINE, n, dim=dim)
        corpus.set_id_base(lo)
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        for c0 in range(0, n, 500_000):
            m = min(500_000, n - c0)
            x = torch.randn((m, dim), device=dev, generator=g)
            corpus.append((x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32))
        gq = torch.Generator(device=dev).manual_seed(99)   # identical queries on every rank
        q = torch.randn((nq, dim), device=dev, generator=gq)
        q = (q / q.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32)
        v = view(corpus, (n, dim), "<f4")
        per = 8 // world if world <= 8 else 1               # 8 planted neighbours per query in total
        prow = torch.randperm(n, device=dev, generator=g)[: nq * max(per, 1)]
        pv = q.repeat_interleave(max(per, 1), dim=0) + 0.03 * torch.randn((nq * max(per, 1), dim), device=dev, generator=g)
        v[prow] = (pv / pv.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32)
        corpus.refresh()
        il, kl = torch.empty((nq, K), dtype=torch.int64, device=dev), torch.empty((nq, K), dtype=torch.float32, device=dev)
        ia, ka = torch.empty((world, nq, K), dtype=torch.int64, device=dev), torch.empty((world, nq, K), dtype=torch.float32, device=dev)
        io, ko = torch.empty_like(il), torch.empty_like(kl)

        def step():
            if world == 1:
                corpus.scan_cosine(q, K, io, ko)
            else:
                corpus.scan_cosine(q, K, il, kl)
                dist.all_gather_into_tensor(ia, il)
                dist.all_gather_into_tensor(ka, kl)
                ctx.merge_topk_f32(ia, ka, world, nq, K, io, ko)
        ctx.profile_begin()
        ms = timed(step, STEPS)
        kms, kfl, kn = ctx.profile_end(_ffi.PROF_COSINE_SCAN)
        assert float(ko[:, 7].min()) > 0.6, "planted neighbours not found"
        line = {"path": "cosine", "metric": "queries/s", "value": nq / ms * 1e3, "ms_per_step": ms, "n_gpus": world, "scaling": "strong",
                "kernel_TFLOPs_per_gpu": kfl / (kms / 1e3) / 1e12,
                "config": {"workload": f"cosine top-{K}, {nq}-query batch over {n_total} x {dim} unit vectors (8 planted/query)", "rows_per_gpu": n}}
    else:
        w = h = 1024
        n = 512 if small else 2048                          # images per rank per step (weak scaling: the batch splits)
        px = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=dev)
        out = torch.zeros((n, 51), dtype=torch.int64, device=dev)
        ms = timed(lambda: ctx.image_hash_uniform(px, n, w, h, out=out), STEPS)
        line = {"path": "image", "metric": "images/s", "value": world * n / ms * 1e3, "ms_per_step": ms, "n_gpus": world, "scaling": "weak",
                "config": {"workload": f"multi bundle on synthetic {w}x{h} RGB images, {n} per GPU per step"}}
if rank == 0:
    print(json.dumps(line), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
