"""The N > 1 path on CPU: two gloo ranks each scan a record-range shard (with the oracle standing in for the
GPU scan), all-gather their per-rank top-k with global ids, run the merge every rank runs, and must
reproduce the single-shard answer byte for byte."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from ucfp_b200.sharding import merge_topk_host, shard_range

N, NQ, K = 40_003, 9, 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(N, rank, world)
    if kind == "hamming":
        codes = oracle.fill_u64(N, 0xC0DE) & np.uint64(0xFFFFF)          # few distinct bits: ties across shards
        q = oracle.fill_u64(NQ, 0xBEEF) & np.uint64(0xFFFFF)
        ids, keys = oracle.hamming_topk(codes[lo:hi], q, K, id_base=lo)
        desc = False
    elif kind == "jaccard":
        sig = oracle.fill_u64(N // 8 * 128, 7).reshape(-1, 128)
        q = oracle.fill_u64(NQ * 128, 8).reshape(NQ, 128)
        for j in range(NQ):
            sig[(j * 611) % len(sig), : 20 + 10 * j] = q[j, : 20 + 10 * j]
            sig[(j * 611 + 3000) % len(sig), : 20 + 10 * j] = q[j, : 20 + 10 * j]
        lo, hi = shard_range(len(sig), rank, world)
        ids, keys = oracle.jaccard_topk(sig[lo:hi], q, K, id_base=lo)
        desc = True
    else:
        rng = np.random.default_rng(5)
        rows = rng.standard_normal((N // 4, 64)).astype(np.float32)
        q = rng.standard_normal((NQ, 64)).astype(np.float32)
        lo, hi = shard_range(len(rows), rank, world)
        ids, keys, _ = oracle.cosine_topk(rows[lo:hi], q, K, id_base=lo, mode=1)
        desc = True
    t_ids = torch.from_numpy(ids.view(np.int64).copy())
    t_keys = torch.from_numpy(keys.view(np.int32).copy())
    g_ids = [torch.empty_like(t_ids) for _ in range(world)]
    g_keys = [torch.empty_like(t_keys) for _ in range(world)]
    dist.all_gather(g_ids, t_ids)
    dist.all_gather(g_keys, t_keys)
    all_ids = torch.stack(g_ids).numpy().view(np.uint64)
    all_keys = torch.stack(g_keys).numpy().view(keys.dtype)
    m_ids, m_keys = merge_topk_host(all_ids, all_keys, K, descending=desc)
    np.save(os.path.join(out_dir, f"{kind}_{rank}_ids.npy"), m_ids)
    np.save(os.path.join(out_dir, f"{kind}_{rank}_keys.npy"), m_keys)
    dist.destroy_process_group()


def _single(kind):
    if kind == "hamming":
        codes = oracle.fill_u64(N, 0xC0DE) & np.uint64(0xFFFFF)
        q = oracle.fill_u64(NQ, 0xBEEF) & np.uint64(0xFFFFF)
        return oracle.hamming_topk(codes, q, K)
    if kind == "jaccard":
        sig = oracle.fill_u64(N // 8 * 128, 7).reshape(-1, 128)
        q = oracle.fill_u64(NQ * 128, 8).reshape(NQ, 128)
        for j in range(NQ):
            sig[(j * 611) % len(sig), : 20 + 10 * j] = q[j, : 20 + 10 * j]
            sig[(j * 611 + 3000) % len(sig), : 20 + 10 * j] = q[j, : 20 + 10 * j]
        return oracle.jaccard_topk(sig, q, K)
    rng = np.random.default_rng(5)
    rows = rng.standard_normal((N // 4, 64)).astype(np.float32)
    q = rng.standard_normal((NQ, 64)).astype(np.float32)
    ids, sc, _ = oracle.cosine_topk(rows, q, K, mode=1)
    return ids, sc


def test_two_rank_gather_and_merge_equals_one_shard(tmp_path):
    world = 2
    for kind in ("hamming", "jaccard", "cosine"):
        mp.spawn(_worker, args=(world, _free_port(), kind, str(tmp_path)), nprocs=world, join=True)
        want_ids, want_keys = _single(kind)
        for r in range(world):   # every rank ends with the same, complete answer
            got_ids = np.load(tmp_path / f"{kind}_{r}_ids.npy")
            got_keys = np.load(tmp_path / f"{kind}_{r}_keys.npy")
            np.testing.assert_array_equal(got_ids, want_ids)
            np.testing.assert_array_equal(got_keys.view(np.uint32), want_keys.view(np.uint32))


def test_shard_ranges_partition_the_corpus():
    for n in (0, 1, 7, 1_000_000_000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
