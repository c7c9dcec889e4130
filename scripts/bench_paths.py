"""Measures every path of the hot path at BASELINE.json's config shapes on ONE B200 and prints one JSON line per
path (committed as profiles/rNN_paths.jsonl).  bench.py stays the contract line (Hamming); this is the evidence
for the other §8 rows: Jaccard (config 3), cosine (config 4), image hashing (configs 1 and 5).

    python scripts/bench_paths.py [jaccard] [cosine] [image] [--small]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (CPU baseline legs only)
from ucfp_b200 import Context, Corpus, _ffi  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
    {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
small = "--small" in sys.argv
want = [a for a in sys.argv[1:] if not a.startswith("--")] or ["jaccard", "cosine", "image"]
ctx = Context(0)
dev = torch.device("cuda", 0)


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def rows_view(corpus, shape, typestr):
    class _A:
        __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (corpus.device_rows_ptr(), False), "version": 2}
    return torch.as_tensor(_A(), device=dev)


if "jaccard" in want:
    n, nq, k = (5_000_000 if small else 50_000_000), 256, 10
    corpus = Corpus(ctx, _ffi.KIND_MINHASH128, n)
    corpus.append_synthetic(0x5EED, 0, n)
    q = oracle.fill_u64(nq * 128, 77).reshape(nq, 128)
    rng = np.random.default_rng(0)
    view = rows_view(corpus, (n, 128), "<i8")
    for lo in range(0, n // 100, 100_000):     # BASELINE config 3: 1 % of rows copy a query's slots with p in {.9,.7,.5}
        m = min(100_000, n // 100 - lo)
        rows = torch.from_numpy(rng.choice(n, m, replace=False)).to(dev)
        base = oracle.fill_u64(m * 128, 99 + lo).reshape(m, 128)
        qi, p = rng.integers(0, nq, m), rng.choice([0.9, 0.7, 0.5], m)
        mask = rng.random((m, 128)) < p[:, None]
        base[mask] = q[qi][mask]
        view[rows] = torch.from_numpy(base.view(np.int64)).to(dev)
    corpus.refresh()
    qd = torch.from_numpy(q.view(np.int64)).to(dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    mt = torch.empty((nq, k), dtype=torch.int32, device=dev)
    ms = timed(lambda: corpus.scan_jaccard(qd, k, ids, mt), 3)
    ctx.profile_begin()
    corpus.scan_jaccard(qd, k, ids, mt)
    kms, kbytes, kn = ctx.profile_end(_ffi.PROF_JACCARD_SCAN)
    fb = ctx.last_scan_fallbacks()
    qh = torch.from_numpy(q.view(np.int64)).pin_memory()
    ih, mh = np.zeros((nq, k), np.uint64), np.zeros((nq, k), np.uint32)
    ms_e = timed(lambda: corpus.scan_jaccard(qh.numpy(), k, ih, mh), 3)
    ms1 = timed(lambda: corpus.scan_jaccard(qd[:1].contiguous(), k, ids[:1], mt[:1]), 5)
    # CPU baseline: oracle on a bounded sample of the same corpus
    sn = 200_000
    sample = view[:sn].cpu().numpy().view(np.uint64)
    t0 = time.perf_counter()
    oi, om = oracle.jaccard_topk(sample, q, k, threads=oracle.host_threads())
    dt = time.perf_counter() - t0
    sub = Corpus(ctx, _ffi.KIND_MINHASH128, sn)
    sub.append(sample)
    gi, gm = sub.scan_jaccard(q, k)
    assert (gi == oi).all() and (gm == om).all(), "jaccard parity failed on the bench sample"
    print(json.dumps({"path": "jaccard", "config": f"MinHash-128 Jaccard top-{k}, {nq}-query batch over {n} synthetic signatures (1 % planted), 1 B200",
                      "metric": "queries/s", "value": nq / ms * 1e3, "ms_per_batch": ms, "fallbacks": fb,
                      "e2e": {"value": nq / ms_e * 1e3, "h2d_bytes_per_step": nq * 1024, "d2h_bytes_per_step": nq * k * 12},
                      "roofline": {"bound": "hbm", "kernel": "jaccard_scan_kernel", "achieved": kbytes / (kms / 1e3) / 1e9, "peak": PEAKS["hbm_gbs"],
                                   "unit": "GB/s", "frac": kbytes / (kms / 1e3) / 1e9 / PEAKS["hbm_gbs"], "launches": kn,
                                   "note": "algorithmic bytes = 1024 B x rows x queries; the scan reads the 128 B/row sketch once per batch, "
                                           "so the batched figure exceeds the DRAM peak by design (ALU pipe is the bound in force)",
                                   "single_query": {"ms": ms1, "alg_GBps": 1024.0 * n / ms1 / 1e6, "sketch_GBps": 128.0 * n / ms1 / 1e6}},
                      "cpu_baseline": {"value": nq / dt * sn / n, "unit": "queries/s", "cores": oracle.host_threads(), "kind": "port",
                                       "sample": f"{nq} queries x first {sn} rows, {dt:.1f} s, scaled linearly to {n} rows"}}), flush=True)
    corpus.close(); sub.close(); del view
    torch.cuda.empty_cache()

if "cosine" in want:
    n, dim, nq, k = (2_000_000 if small else 20_000_000), 512, 1024, 10
    corpus = Corpus(ctx, _ffi.KIND_COSINE, n, dim=dim)
    g = torch.Generator(device=dev).manual_seed(1)
    for lo in range(0, n, 1_000_000):          # BASELINE config 4: unit-norm rows with bf16-representable values
        m = min(1_000_000, n - lo)
        x = torch.randn((m, dim), device=dev, generator=g)
        x = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32)
        corpus.append(x)
    q = torch.randn((nq, dim), device=dev, generator=g)
    q = (q / q.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32)
    view = rows_view(corpus, (n, dim), "<f4")
    prow = torch.randint(0, n, (nq * 8,), device=dev, generator=g)
    noise = torch.randn((nq * 8, dim), device=dev, generator=g) * 0.03
    pv = q.repeat_interleave(8, dim=0) + noise
    view[prow] = (pv / pv.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32)   # 8 planted neighbours per query
    corpus.refresh()
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ms = timed(lambda: corpus.scan_cosine(q, k, ids, sc), 3)
    ctx.profile_begin()
    corpus.scan_cosine(q, k, ids, sc)
    kms, kflop, kn = ctx.profile_end(_ffi.PROF_COSINE_SCAN)
    fb = ctx.last_scan_fallbacks()
    qh = q.cpu().pin_memory()
    ih, sh = np.zeros((nq, k), np.uint64), np.zeros((nq, k), np.float32)
    ms_e = timed(lambda: corpus.scan_cosine(qh.numpy(), k, ih, sh), 3)
    sn, sq = 200_000, 64
    sample = view[:sn].cpu().numpy()
    t0 = time.perf_counter()
    oi, osc, _ = oracle.cosine_topk(sample, qh.numpy()[:sq], k, mode=1, threads=oracle.host_threads())
    dt = time.perf_counter() - t0
    sub = Corpus(ctx, _ffi.KIND_COSINE, sn, dim=dim)
    sub.append(sample)
    gi, gs = sub.scan_cosine(qh.numpy()[:sq].copy(), k)
    assert (gi == oi).all() and (gs.view(np.uint32) == osc.view(np.uint32)).all(), "cosine parity failed on the bench sample"
    tf = kflop / (kms / 1e3) / 1e12
    print(json.dumps({"path": "cosine", "config": f"cosine top-{k}, {nq}-query batch over {n} x {dim} synthetic unit vectors (bf16-representable, 8 planted/query), 1 B200",
                      "metric": "queries/s", "value": nq / ms * 1e3, "ms_per_batch": ms, "fallbacks": fb,
                      "e2e": {"value": nq / ms_e * 1e3, "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * k * 12},
                      "roofline": {"bound": "tensor", "kernel": "cosine_coarse_kernel", "achieved": tf, "peak": PEAKS["bf16_tflops_sustained"] ,
                                   "unit": "TFLOP/s", "frac": tf / PEAKS["bf16_tflops_sustained"], "launches": kn,
                                   "note": "2 x rows x dim x queries flop per launch (bf16 tcgen05, f32 accumulate); sustained cuBLAS peak"},
                      "cpu_baseline": {"value": sq / dt * sn / n, "unit": "queries/s", "cores": oracle.host_threads(), "kind": "port",
                                       "sample": f"{sq} queries x first {sn} rows, {dt:.1f} s, scaled linearly to {n} rows"}}), flush=True)
    corpus.close(); sub.close(); del view
    torch.cuda.empty_cache()

if "image" in want:
    for (w, h, n, cfg) in ((256, 256, 10_000, "config 1: multi bundle on 10 000 synthetic 256x256 RGB images"),
                           (1024, 1024, 2048 if small else 8192, "config 5 chunk: multi bundle on synthetic 1024x1024 RGB images")):
        words = w * h * 3 // 8
        px = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
        tmp = Corpus(ctx, _ffi.KIND_HAMMING64, 1 << 24)
        for lo in range(0, n * words, 1 << 24):      # device generator: bytes = LE view of splitmix64(seed, i)
            m = min(1 << 24, n * words - lo)
            tmp.clear(); tmp.append_synthetic(0x1316, lo, m)
            src = rows_view(tmp, (m,), "<i8")
            px.view(-1).view(torch.int64)[lo:lo + m] = src
        tmp.close()
        out = torch.zeros((n, 51), dtype=torch.int64, device=dev)
        ms = timed(lambda: ctx.image_hash_uniform(px, n, w, h, out=out), 5)
        nh = min(n, 2048 if w == 256 else 128)
        pxh = px[:nh].cpu().pin_memory()
        outh = np.zeros((nh, 51), np.uint64)
        ms_e = timed(lambda: ctx.image_hash_uniform(pxh.numpy(), nh, w, h, out=outh), 3)
        sn = 256 if w == 256 else 32
        t0 = time.perf_counter()
        want_words = oracle.image_multihash_batch(pxh.numpy()[:sn], threads=oracle.host_threads())
        dt = time.perf_counter() - t0
        assert (out[:sn].cpu().numpy().view(np.uint64) == want_words).all(), "image parity failed on the bench sample"
        gbps = n * (3.0 * w * h + 408) / (ms / 1e3) / 1e9
        print(json.dumps({"path": "image", "config": cfg + f" ({n} per step), 1 B200", "metric": "images/s", "value": n / ms * 1e3, "ms_per_batch": ms,
                          "e2e": {"value": nh / ms_e * 1e3, "h2d_bytes_per_step": nh * 3 * w * h, "d2h_bytes_per_step": nh * 408},
                          "roofline": {"bound": "hbm", "kernel": "image_stream_kernel", "achieved": gbps, "peak": PEAKS["hbm_gbs"], "unit": "GB/s",
                                       "frac": gbps / PEAKS["hbm_gbs"], "note": "3*w*h + 408 B per image"},
                          "cpu_baseline": {"value": sn / dt, "unit": "images/s", "cores": oracle.host_threads(), "kind": "port",
                                           "sample": f"{sn} images, {dt:.2f} s"}}), flush=True)
        del px, out
        torch.cuda.empty_cache()
