"""Robustness of the threshold-filter design on SKEWED corpora: every throughput figure of the
headline bench is measured on uniform random codes, but real perceptual-hash corpora are clustered -- near-duplicate families,
all-zero / all-one hashes of flat images -- and real MinHash corpora hold exact duplicates.  One JSON line per case:
queries/s, queries recomputed by the exact fallback, the longest candidate list (capacity 4096), and a parity check of a
sample of the queries against the CPU oracle over the whole corpus.

    python scripts/bench_robustness.py [--rows 1e9] [--small]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (parity checker only)
from ucfp_b200 import Context, Corpus, _ffi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=float, default=1e9)
ap.add_argument("--small", action="store_true")
args = ap.parse_args()
N = int(2e7 if args.small else args.rows)
ctx = Context(0)
dev = torch.device("cuda", 0)
K, NQ = 10, 1024
U64 = np.uint64


def view_of(corpus, shape):
    class _A:
        __cuda_array_interface__ = {"shape": shape, "typestr": "<i8", "data": (corpus.device_rows_ptr(), False), "version": 2}
    return torch.as_tensor(_A(), device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def flips(gen, n, max_bits):
    """n random u64 masks with 0..max_bits bits set (as int64 tensors)"""
    m = torch.zeros(n, dtype=torch.int64, device=dev)
    for _ in range(max_bits):
        bit = torch.randint(0, 64, (n,), device=dev, generator=gen)
        on = torch.rand(n, device=dev, generator=gen) < 0.75
        m ^= torch.where(on, torch.ones_like(m) << bit, torch.zeros_like(m))
    return m


# ---- case 1: clustered Hamming corpus ----------------------------------------------------------------------------------
# 1 % of the rows lie within 4 bit flips of one of 1000 centres (near-duplicate families), 0.1 % are the all-zero code (flat
# images), the rest is uniform.  Queries: 512 members of families, 128 centres, 128 zeros / all-ones, 256 uniform codes.
def clustered(explicit_ids):
    gen = torch.Generator(device=dev).manual_seed(7)
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, N)
    corpus.append_synthetic(0xC0DE, 0, N)
    v = view_of(corpus, (N,))
    centres = torch.randint(-2**62, 2**62, (1000,), device=dev, generator=gen, dtype=torch.int64)
    n_cl, n_zero = N // 100, N // 1000
    rows = torch.randint(0, N, (n_cl,), device=dev, generator=gen)
    fam = torch.randint(0, 1000, (n_cl,), device=dev, generator=gen)
    v[rows] = centres[fam] ^ flips(gen, n_cl, 4)
    v[torch.randint(0, N, (n_zero,), device=dev, generator=gen)] = 0
    ids = None
    if explicit_ids:   # a random permutation-like id column: ties are no longer broken in scan order
        ids = (torch.arange(N, device=dev, dtype=torch.int64) * 0x9E3779B97F4A7C15 + 12345) & 0x7FFFFFFFFFFFFFFF
        # explicit ids need an all-explicit corpus: rebuild it from the resident rows
        rows_copy = v.clone()
        corpus.close()
        corpus = Corpus(ctx, _ffi.KIND_HAMMING64, N)
        corpus.append(rows_copy, ids)
        v = view_of(corpus, (N,))
        del rows_copy
    else:
        corpus.refresh()
    qf = torch.randint(0, 1000, (512,), device=dev, generator=gen)
    q = torch.cat([centres[qf] ^ flips(gen, 512, 4), centres[:128], torch.zeros(64, dtype=torch.int64, device=dev),
                   torch.full((64,), -1, dtype=torch.int64, device=dev), torch.randint(-2**62, 2**62, (256,), device=dev, generator=gen, dtype=torch.int64)])
    io, do = torch.empty((NQ, K), dtype=torch.int64, device=dev), torch.empty((NQ, K), dtype=torch.int32, device=dev)
    ms = timed(lambda: corpus.scan_hamming(q, K, io, do))
    fb, fill = ctx.last_scan_stats()
    # parity: the oracle over the whole corpus for 32 of the queries (4 of each kind at least)
    sel = np.concatenate([np.arange(0, 512, 32), np.arange(512, 640, 16), np.arange(640, 768, 32), np.arange(768, 1024, 64)])
    codes_h = v.cpu().numpy().view(U64)
    ids_h = ids.cpu().numpy().view(U64) if ids is not None else None
    t0 = time.perf_counter()
    oi, od = oracle.hamming_topk(codes_h, np.ascontiguousarray(q.cpu().numpy().view(U64)[sel]), K, ids=ids_h, threads=oracle.host_threads())
    ok = bool((io.cpu().numpy().view(U64)[sel] == oi).all() and (do.cpu().numpy().view(np.uint32)[sel] == od).all())
    print(json.dumps({"case": "hamming_clustered" + ("_explicit_ids" if explicit_ids else ""), "rows": N, "queries": NQ, "k": K,
                      "corpus": "1 % within 4 flips of 1000 centres, 0.1 % all-zero, rest uniform",
                      "queries_per_s": NQ / ms * 1e3, "ms_per_batch": ms, "fallbacks": fb, "max_list_fill": fill, "list_capacity": 4096,
                      "parity_check": {"queries": len(sel), "ok": ok, "oracle_seconds": round(time.perf_counter() - t0, 1)},
                      "kth_dist_family_queries_median": float(np.median(do.cpu().numpy()[:512, -1]))}), flush=True)
    corpus.close()
    return ok


# ---- case 2: duplicate-heavy MinHash corpus ----------------------------------------------------------------------------------
def jaccard_duplicates():
    n = 2_000_000 if args.small else 20_000_000
    gen = torch.Generator(device=dev).manual_seed(9)
    corpus = Corpus(ctx, _ffi.KIND_MINHASH128, n)
    corpus.append_synthetic(0x5EED, 0, n)
    v = view_of(corpus, (n, 128))
    protos = v[:100].clone()                                   # 100 documents ...
    dup_rows = torch.randint(100, n, (n // 5,), device=dev, generator=gen)
    for a in range(0, len(dup_rows), 200_000):                 # ... re-ingested verbatim as 20 % of the corpus
        r = dup_rows[a:a + 200_000]
        v[r] = protos[torch.randint(0, 100, (len(r),), device=dev, generator=gen)]
    corpus.refresh()
    nq = 256
    q = torch.cat([protos, v[torch.randint(0, n, (156,), device=dev, generator=gen)]])
    io, mo = torch.empty((nq, K), dtype=torch.int64, device=dev), torch.empty((nq, K), dtype=torch.int32, device=dev)
    ms = timed(lambda: corpus.scan_jaccard(q, K, io, mo), 3)
    fb, fill = ctx.last_scan_stats()
    sel = np.arange(0, nq, 16)
    sn = min(n, 4_000_000)                                    # oracle over a prefix sub-corpus (1 KiB rows: host memory)
    sub = Corpus(ctx, _ffi.KIND_MINHASH128, sn)
    sub.append(v[:sn].contiguous())
    gi, gm = sub.scan_jaccard(q[sel].contiguous(), K)
    oi, om = oracle.jaccard_topk(v[:sn].cpu().numpy().view(U64), q[sel].cpu().numpy().view(U64), K, threads=oracle.host_threads())
    ok = bool((gi.cpu().numpy().view(U64) == oi).all() and (gm.cpu().numpy().view(np.uint32) == om).all())
    print(json.dumps({"case": "jaccard_duplicates", "rows": n, "queries": nq, "k": K, "corpus": "20 % of the rows are verbatim copies of 100 signatures",
                      "queries_per_s": nq / ms * 1e3, "ms_per_batch": ms, "fallbacks": fb, "max_list_fill": fill, "list_capacity": 4096,
                      "parity_check": {"queries": len(sel), "rows": sn, "ok": ok}, "all_duplicate_queries_hit_128": bool((mo[:100] == 128).all().item())}), flush=True)
    sub.close(); corpus.close()
    return ok


good = clustered(False)
good = clustered(True) and good
good = jaccard_duplicates() and good
sys.exit(0 if good else 3)
