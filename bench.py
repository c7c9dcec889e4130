#!/usr/bin/env python
"""bench.py -- contract benchmark for the UCFP B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric, "Hamming top-10 queries/s over 1B 64-bit hashes"): one STEP is one
1024-query batch scanned against a corpus of 1 B synthetic 64-bit codes with planted neighbours
(BASELINE config 2's construction at the metric's corpus size), k = 10.  The corpus is sharded by
record range over the N ranks (strong scaling: 1 B rows in total at every N); each rank scans its
slice, NCCL all-gathers the per-rank top-k candidates and every rank runs the same deterministic merge.

One JSON line on stdout (rank 0).  `value` = queries/s with queries and results resident in HBM;
`e2e` = the same through the C ABI with pinned HOST query/result buffers (H2D + D2H inside the timed
region); `roofline` = the dominant kernel (hamming_mma_scan_kernel, the int8 tensor-core form of the scan
that batches of >= 64 queries use) timed live with CUDA events inside the library, with the HBM-bound
one-query-per-pass regime of hamming_scan_kernel beside it (`streaming`); `cpu_baseline` = the CPU oracle
on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hamming_top10_queries_per_s_over_1B_codes"
UNIT = "queries/s"
SEED_CORPUS, SEED_QUERY, SEED_PLANT = 0xC0DE, 0xBEEF, 0xFACE
K = 10


# ---------------------------------------------------------------- synthetic data (numpy, no oracle) --
def _mix64(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """docs/HASH_SPEC.md section 8: mix64(seed ^ (index + 1) * golden)."""
    with np.errstate(over="ignore"):
        return _mix64(np.uint64(seed) ^ ((idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)))


def make_queries(nq: int) -> np.ndarray:
    return splitmix64(SEED_QUERY, np.arange(nq, dtype=np.uint64))


def planted(nq: int, n_total: int):
    """12 neighbours per query at Hamming distance 0..11 at pseudo-random rows (BASELINE config 2)."""
    q = make_queries(nq)
    j = np.arange(nq * 12, dtype=np.uint64)
    rows = splitmix64(SEED_PLANT, j) % np.uint64(n_total)
    d = (j % np.uint64(12)).astype(np.uint64)
    codes = np.repeat(q, 12) ^ ((np.uint64(1) << d) - np.uint64(1))
    uniq, first = np.unique(rows, return_index=True)  # colliding rows keep the first write
    return uniq, codes[first]


# ---------------------------------------------------------------- clocks ------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.rows, self._stop, self._t = gpu_index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu_index)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bench_config(n_total: int, nq: int, world: int) -> dict:
    """The `config` object of the JSON line.  Both arms (ours and --impl reference) print exactly this dict."""
    return {"workload": f"hamming top-{K}, {nq}-query batch over {n_total} synthetic 64-bit codes with planted neighbours",
            "k": K, "queries": nq, "codes": n_total, "codes_per_gpu": n_total * 1 // world if world else n_total,
            "parallelism": f"record-range shards x{world}, NCCL all-gather of top-k + merge",
            "l2": "inputs larger than L2 (corpus slice >= 1 GB per GPU vs 126 MB L2), no flush needed"}


# ---------------------------------------------------------------- parity (oracle as the checker) ------
def oracle_topk_full_corpus(n_total: int, nq: int, sel: np.ndarray, chunk: int = 100_000_000):
    """CPU oracle (oracle/, the executable form of docs/HASH_SPEC.md section 6) over ALL n_total rows of the bench
    corpus -- regenerated on the host chunk by chunk with the same PRNG and the same planting -- for the queries
    `sel`.  Returns (ids [len(sel), K] u64, dist [len(sel), K] u32).  Never inside a timed region."""
    import oracle
    threads = oracle.host_threads()
    q = make_queries(nq)[sel]
    prow, pcode = planted(nq, n_total)
    best_i = np.full((len(sel), K), np.uint64(2**64 - 1), dtype=np.uint64)
    best_d = np.full((len(sel), K), np.uint32(2**32 - 1), dtype=np.uint32)
    for lo in range(0, n_total, chunk):
        m = min(chunk, n_total - lo)
        codes = oracle.fill_u64(m, SEED_CORPUS, start=lo)
        inside = (prow >= np.uint64(lo)) & (prow < np.uint64(lo + m))
        codes[(prow[inside] - np.uint64(lo)).astype(np.int64)] = pcode[inside]
        oi, od = oracle.hamming_topk(codes, q, K, id_base=lo, threads=threads)
        # merge under (dist asc, id asc); sentinels (id = 2^64 - 1, dist = 2^32 - 1) sort last by construction
        ci, cd = np.concatenate([best_i, oi], axis=1), np.concatenate([best_d, od], axis=1)
        for r in range(len(sel)):
            order = np.lexsort((ci[r], cd[r]))[:K]
            best_i[r], best_d[r] = ci[r][order], cd[r][order]
        del codes
    return best_i, best_d


# ---------------------------------------------------------------- CPU arm -----------------------------
def cpu_arm(n_total: int, nq: int, steps: int, warmup: int, sample_rows: int, sample_queries: int):
    """The reference's CPU path for this metric.  The Rust reference has no Hamming scan at all and cannot
    be built here (no cargo), so this is the CPU restatement (oracle/, kind="port") on all host threads.
    Each step scans `sample_queries` queries over the first `sample_rows` rows of the same synthetic corpus;
    queries/s over 1 B rows = measured queries/s x sample_rows / 1e9 (brute force is linear in rows)."""
    import oracle
    threads = oracle.host_threads()
    codes = oracle.fill_u64(sample_rows, SEED_CORPUS)
    q = make_queries(nq)[:sample_queries]
    for _ in range(min(warmup, 1)):
        oracle.hamming_topk(codes[: sample_rows // 8], q, K, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.hamming_topk(codes, q, K, threads=threads)
    dt = time.perf_counter() - t0
    qps_sample = steps * sample_queries / dt
    value = qps_sample * sample_rows / n_total
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample_queries} queries x first {sample_rows} rows of the corpus per step, {steps} step(s), "
                      f"{dt:.1f} s; scaled linearly to {n_total} rows"}, dt / steps * 1e3


# ---------------------------------------------------------------- secondary paths (configs 3 and 4) --------
def _rows_view(torch, corpus, shape, typestr, dev):
    class _A:
        __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (corpus.device_rows_ptr(), False), "version": 2}
    return torch.as_tensor(_A(), device=dev)


def _crc(ids, keys) -> int:
    import zlib
    return zlib.crc32(np.ascontiguousarray(keys).tobytes(), zlib.crc32(np.ascontiguousarray(ids).tobytes()))


def secondary_jpeg(ctx, rank, world, check):
    """SURVEY 8f N3 / the reference's only image bench shape (benches/end_to_end.rs:40-53 starts from ENCODED bytes): images/s from
    JPEG bitstreams in host memory -- nvJPEG decode on the device + multi-hash in ONE C-ABI call -- next to the host doing the
    same job (Pillow = libjpeg-turbo decode + the oracle's hash, one thread, a bounded sample).  Wall clock: the call is
    synchronous and starts from host bytes."""
    import io
    import time
    try:
        from PIL import Image
    except Exception as e:  # no encoder on this box: nothing to feed the path with
        return {"unavailable": f"Pillow missing: {e}"}
    w = h = 1024
    blobs = []
    for seed in range(8):   # smooth synthetic scenes (JPEG of noise is meaningless), quality 85, 4:2:0
        rng = np.random.default_rng(100 + seed)
        y, x = np.mgrid[0:h, 0:w].astype(np.float32)
        img = np.zeros((h, w, 3), np.float32)
        for c in range(3):
            for _ in range(6):
                fx, fy = rng.uniform(0.002, 0.05, 2)
                img[..., c] += rng.uniform(20, 60) * np.sin(fx * x + fy * y + rng.uniform(0, 6.28))
        buf = io.BytesIO()
        Image.fromarray(np.clip(img + 128, 0, 255).astype(np.uint8)).save(buf, format="JPEG", quality=85, subsampling=2)
        blobs.append(buf.getvalue())
    batch = blobs * 32                                    # 256 images per call
    ctx.image_hash_jpeg_batch(batch[:16])
    ctx.image_hash_jpeg_batch(batch)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        words, status, _ = ctx.image_hash_jpeg_batch(batch)
    dt = (time.perf_counter() - t0) / reps
    out = {"metric": "images_hashed_per_s_from_jpeg_bytes", "value": world * len(batch) / dt, "unit": "images/s",
           "images_per_gpu_per_step": len(batch), "encoded_mb_per_step": sum(len(b) for b in batch) / 1e6, "shape": "1024x1024, q85, 4:2:0",
           "decoded_on_device": int((status == 0).sum()), "timing": "wall clock around the synchronous C-ABI call, host bytes in, host hashes out"}
    if rank == 0 and check:
        import oracle
        w2, st2, _, px = ctx.image_hash_jpeg_batch(blobs[:4], want_pixels=True)
        ok = bool((st2 == 0).all()) and all((w2[i] == oracle.image_multihash(px[i])).all() for i in range(4))
        out["parity_check"] = {"images": 4, "ok": ok, "against": "oracle/ over the pixels the device decoded (JPEG decoders are not bit-identical to each other)"}
        t0 = time.perf_counter()
        n_cpu = 8
        for b in blobs[:n_cpu]:
            oracle.image_multihash(np.asarray(Image.open(io.BytesIO(b)).convert("RGB")))
        out["cpu_decode_and_hash"] = {"value": n_cpu / (time.perf_counter() - t0), "unit": "images/s", "cores": 1,
                                      "kind": "port", "sample": f"{n_cpu} images: Pillow (libjpeg-turbo) decode + oracle multi-hash, one thread"}
    return out


def secondary_jaccard(torch, ctx, group, rank, world, dev, timed, peak, small):
    """BASELINE configs[2]: MinHash-128 Jaccard top-10 over 50 M synthetic signatures (1 % of the rows copy a query's slots
    with p in {.9,.7,.5}), 256-query batch, record-range shards over the N ranks.  Row contents depend on the GLOBAL row
    index only, so `result_crc` must be the same at every N."""
    from ucfp_b200 import Corpus, _ffi
    n_total, nq, k = (5_000_000 if small else 50_000_000), 256, 10
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    corpus = Corpus(ctx, _ffi.KIND_MINHASH128, hi - lo)
    corpus.set_id_base(lo)
    corpus.append_synthetic(0x5EED, lo, hi - lo)
    q = splitmix64(77, np.arange(nq * 128, dtype=np.uint64)).reshape(nq, 128)
    prow = np.unique(splitmix64(0x9A, np.arange(n_total // 100, dtype=np.uint64)) % np.uint64(n_total))
    mine = prow[(prow >= lo) & (prow < hi)]
    view = _rows_view(torch, corpus, (hi - lo, 128), "<i8", dev)
    qd = torch.from_numpy(q.view(np.int64)).to(dev)
    for a in range(0, len(mine), 100_000):
        rows = mine[a:a + 100_000]
        h = splitmix64(0x9B, rows)
        qi = torch.from_numpy((h % np.uint64(nq)).astype(np.int64)).to(dev)
        p = np.array([0.9, 0.7, 0.5])[((h >> np.uint64(20)) % np.uint64(3)).astype(np.int64)]
        u = (splitmix64(0x9C, (rows[:, None] * np.uint64(128) + np.arange(128, dtype=np.uint64)[None, :]).reshape(-1)) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
        mask = torch.from_numpy(u.reshape(-1, 128) < p[:, None]).to(dev)
        loc = torch.from_numpy((rows - np.uint64(lo)).astype(np.int64)).to(dev)
        view[loc] = torch.where(mask, qd[qi], view[loc])
    corpus.refresh()
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    m = torch.empty((nq, k), dtype=torch.int32, device=dev)
    step = (lambda: corpus.scan_jaccard(qd, k, ids, m)) if world == 1 else (lambda: group.scan_jaccard([corpus], qd, k, ids, m))
    for _ in range(2):
        step()
    ctx.profile_begin()
    ms = timed(step, 3) / 3
    kms, kbytes, kn = ctx.profile_end(_ffi.PROF_JACCARD_SCAN)
    gi, gm = ids.cpu().numpy().view(np.uint64), m.cpu().numpy().view(np.uint32)
    out = {"metric": "jaccard_top10_queries_per_s_over_50M_signatures" if not small else "jaccard_top10_queries_per_s_over_5M_signatures",
           "value": nq / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "queries": nq, "signatures": n_total, "result_crc": _crc(gi, gm),
           "best_matches_min": int(gm[:, 0].min()),
           "roofline": {"bound": "hbm", "kernel": "jaccard_scan_kernel", "achieved": kbytes / (kms / 1e3) / 1e9 if kms else None, "peak": peak,
                        "unit": "GB/s", "frac": kbytes / (kms / 1e3) / 1e9 / peak if kms else None, "launches": kn,
                        "note": "algorithmic bytes = 1024 B x rows x queries per launch (this rank's shard); the scan streams the 128 B/row "
                                "sketch once per batch, so the batched figure exceeds the DRAM peak by design"}}
    # parity on a bounded sample of THIS data (full-size parity lives in tests/test_configs_gpu.py): first rows of rank 0's shard
    if rank == 0:
        import oracle
        sn = 100_000
        sample = view[:sn].cpu().numpy().view(np.uint64)
        sub = Corpus(ctx, _ffi.KIND_MINHASH128, sn)
        sub.append(sample)
        si, sm = sub.scan_jaccard(q[:32].copy(), k)
        oi, om = oracle.jaccard_topk(sample, q[:32].copy(), k, threads=oracle.host_threads())
        out["parity_check"] = {"rows": sn, "queries": 32, "ok": bool((si == oi).all() and (sm == om).all()), "against": "oracle/ on the first rows of the shard"}
        sub.close()
    del view
    corpus.close()
    torch.cuda.empty_cache()
    return out


def secondary_cosine(torch, ctx, group, rank, world, dev, timed, bf16_peak, small):
    """BASELINE configs[3]: cosine top-10 over 20 M x 512 unit vectors with bf16-representable values, 8 planted neighbours
    per query, 1024-query batch, record-range shards.  Rows are generated per global 500 K-row block with a fixed seed."""
    from ucfp_b200 import Corpus, _ffi
    n_total, dim, nq, k, blk = (2_000_000 if small else 20_000_000), 512, 1024, 10, 500_000
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    corpus = Corpus(ctx, _ffi.KIND_COSINE, hi - lo, dim=dim)
    corpus.set_id_base(lo)
    g = torch.Generator(device=dev)

    def unit(x):
        return (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32)

    for b in range(lo // blk, (hi + blk - 1) // blk):
        g.manual_seed(1000 + b)
        x = unit(torch.randn((blk, dim), device=dev, generator=g))
        a, e = max(lo, b * blk), min(hi, (b + 1) * blk)
        corpus.append(x[a - b * blk: e - b * blk].contiguous())
    g.manual_seed(7)
    q = unit(torch.randn((nq, dim), device=dev, generator=g))
    noise = torch.randn((nq * 8, dim), device=dev, generator=g) * 0.03
    planted_rows = splitmix64(0xC05, np.arange(nq * 8, dtype=np.uint64)) % np.uint64(n_total)
    uniq, first = np.unique(planted_rows, return_index=True)
    sel = first[(uniq >= lo) & (uniq < hi)]
    if len(sel):
        view = _rows_view(torch, corpus, (hi - lo, dim), "<f4", dev)
        pv = unit(q.repeat_interleave(8, dim=0) + noise)
        seli = torch.from_numpy(sel.astype(np.int64)).to(dev)
        view[torch.from_numpy((planted_rows[sel] - np.uint64(lo)).astype(np.int64)).to(dev)] = pv[seli]
        del view
        corpus.refresh()
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    step = (lambda: corpus.scan_cosine(q, k, ids, sc)) if world == 1 else (lambda: group.scan_cosine([corpus], q, k, ids, sc))
    for _ in range(2):
        step()
    ctx.profile_begin()
    ms = timed(step, 3) / 3
    kms, kflop, kn = ctx.profile_end(_ffi.PROF_COSINE_SCAN)
    gi, gs = ids.cpu().numpy().view(np.uint64), sc.cpu().numpy()
    tf = kflop / (kms / 1e3) / 1e12 if kms else None
    out = {"metric": f"cosine_top10_queries_per_s_over_{n_total // 1_000_000}M_x_512", "value": nq / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
           "queries": nq, "vectors": n_total, "result_crc": _crc(gi, gs.view(np.uint32)), "best_score_min": float(gs[:, 0].min()),
           "roofline": {"bound": "tensor", "kernel": "cosine_coarse_kernel", "achieved": tf, "peak": bf16_peak, "unit": "TFLOP/s",
                        "frac": tf / bf16_peak if tf else None, "launches": kn,
                        "note": "2 x rows x dim x queries flop per launch (this rank's shard; bf16 tcgen05, f32 accumulate)"}}
    if rank == 0:
        import oracle
        sn, sq = 100_000, 32
        view = _rows_view(torch, corpus, (hi - lo, dim), "<f4", dev)
        sample = view[:sn].cpu().numpy()
        del view
        qh = q[:sq].cpu().numpy()
        sub = Corpus(ctx, _ffi.KIND_COSINE, sn, dim=dim)
        sub.append(sample)
        si, ss = sub.scan_cosine(qh.copy(), k)
        oi, osc, _ = oracle.cosine_topk(sample, qh, k, mode=1, threads=oracle.host_threads())
        out["parity_check"] = {"rows": sn, "queries": sq, "ok": bool((si == oi).all() and (ss.view(np.uint32) == osc.view(np.uint32)).all()),
                               "against": "oracle/ (restatement of src/index/embedded/mod.rs:268-360) on the first rows of the shard, bit-exact f32 scores"}
        sub.close()
    corpus.close()
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------- main --------------------------------
def committed_tensor_pipe_active():
    """sm__pipe_tensor_cycles_active (fraction) of the committed ncu capture of the tensor scan; None when the file is absent."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_hamming_mma_q1024_ncu.txt")
    try:
        with open(path) as f:
            for line in f:
                if line.startswith("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):
                    return round(float(line.split()[-1]) / 100.0, 4)
    except OSError:
        pass
    return None


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--codes", type=float, default=1e9, help="total corpus rows over all ranks")
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--cpu-sample-rows", type=float, default=2e8)
    ap.add_argument("--cpu-sample-queries", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-images", action="store_true", help="skip the secondary images-hashed/s measurement")
    ap.add_argument("--no-paths", action="store_true", help="skip the secondary Jaccard (config 3) and cosine (config 4) measurements")
    ap.add_argument("--small-paths", action="store_true", help="secondary Jaccard / cosine at a tenth of the config sizes (quick runs)")
    ap.add_argument("--parity-queries", type=int, default=64,
                    help="queries of the batch checked against the CPU oracle over the FULL corpus (untimed; 0 = skip)")
    args = ap.parse_args()

    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are sent to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj) -> None:
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_total, nq = int(args.codes), args.queries

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, ms = cpu_arm(n_total, nq, max(args.steps, 1), args.warmup, int(args.cpu_sample_rows), args.cpu_sample_queries)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": bench_config(n_total, nq, max(args.gpus, 1)),
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from ucfp_b200 import Context, Corpus, Group, _ffi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ucfp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ctx = Context(local_rank)  # shares torch's current stream
    # N > 1: one process per GPU.  torch.distributed is the rendezvous only (it ships the NCCL id and keeps the barriers);
    # the data path -- bound exchange, all-gather of packed top-k records, merge -- runs inside the library's group scan.
    group = None
    if world > 1:
        uid = [Group.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        group = Group.join(ctx, uid[0], rank, world)
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    shard = hi - lo
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, shard)
    corpus.set_id_base(lo)
    corpus.append_synthetic(SEED_CORPUS, lo, shard)

    # plant the neighbours that fall into this shard (written straight into the resident rows)
    rows, codes = planted(nq, n_total)
    mine = (rows >= lo) & (rows < hi)

    class _Rows:
        __cuda_array_interface__ = {"shape": (shard,), "typestr": "<i8", "data": (corpus.device_rows_ptr(), False), "version": 2}

    view = torch.as_tensor(_Rows(), device=dev)
    if mine.any():
        view[torch.from_numpy((rows[mine] - np.uint64(lo)).astype(np.int64)).to(dev)] = \
            torch.from_numpy(codes[mine].view(np.int64)).to(dev)
        corpus.refresh()   # rows were written in place: rebuild the tensor scan's operand rows

    queries = make_queries(nq)
    q_host = torch.from_numpy(queries.view(np.int64)).pin_memory()
    q_dev = q_host.to(dev)
    ids_out = torch.empty((nq, K), dtype=torch.int64, device=dev)
    dist_out = torch.empty((nq, K), dtype=torch.int32, device=dev)
    ids_host = torch.empty((nq, K), dtype=torch.int64).pin_memory()
    dist_host = torch.empty((nq, K), dtype=torch.int32).pin_memory()

    def step_device():
        """queries and results resident in HBM"""
        if world == 1:
            corpus.scan_hamming(q_dev, K, ids_out, dist_out)
        else:
            group.scan_hamming([corpus], q_dev, K, ids_out, dist_out)

    def step_e2e():
        """pinned host queries in, pinned host results out, through the C ABI's host-buffer path"""
        if world == 1:
            corpus.scan_hamming(q_host.numpy(), K, ids_host.numpy(), dist_host.numpy())
        else:
            group.scan_hamming([corpus], q_host.numpy(), K, ids_host.numpy(), dist_host.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
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
        return float(ms.item())

    # nvidia-smi answers in ~50-100 ms, a timed region of K steps can be shorter than that: the sampler runs from the
    # warm-up through the device-timed and the end-to-end timed regions (the same step under the same load throughout)
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    launches0 = ctx.kernel_launches
    ctx.profile_begin()
    total_ms = timed(step_device, args.steps)
    t_ms, t_ops, t_n = ctx.profile_read(_ffi.PROF_HAMMING_TENSOR)
    k_ms, k_bytes, k_n = ctx.profile_end(_ffi.PROF_HAMMING_SCAN)
    launches = ctx.kernel_launches - launches0
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(lt)
    value = args.steps * nq / (total_ms / 1e3)

    # the result that was just timed (all N ranks hold the same merged answer); checked against the oracle below
    got_ids_dev = ids_out.cpu().numpy().view(np.uint64).copy()
    got_dist_dev = dist_out.cpu().numpy().view(np.uint32).copy()
    if not (got_dist_dev[:, 0] == 0).all():
        raise SystemExit("bench self-check failed: planted exact duplicates not found")

    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    e2e_value = args.steps * nq / (e2e_ms / 1e3)
    got_ids_e2e = ids_host.numpy().view(np.uint64).copy()
    got_dist_e2e = dist_host.numpy().view(np.uint32).copy()
    if len(clocks.rows) < 3:   # very short runs: keep the same scan going until the sampler has seen it.  LOCAL work only --
        t_end = time.time() + 1.0   # ranks may disagree about needing this, so no collective may run in here
        scratch_i, scratch_d = torch.empty_like(ids_out), torch.empty_like(dist_out)
        while time.time() < t_end:
            corpus.scan_hamming(q_dev, K, scratch_i, scratch_d)
        torch.cuda.synchronize()
    clocks.__exit__(None, None, None)

    # the same kernel in its HBM-bound regime: one and two queries per corpus pass
    streaming = {}
    for b in (1, 2):
        qb = q_dev[:b].contiguous()
        io, do = torch.empty((b, K), dtype=torch.int64, device=dev), torch.empty((b, K), dtype=torch.int32, device=dev)
        for _ in range(3):
            corpus.scan_hamming(qb, K, io, do)
        reps = 20
        ctx.profile_begin()
        ms_b = timed(lambda: corpus.scan_hamming(qb, K, io, do), reps)
        km, kb, kn = ctx.profile_end(_ffi.PROF_HAMMING_SCAN)
        streaming[f"q{b}"] = {"ms_per_pass": ms_b / reps, "kernel_GBps": kb / (km / 1e3) / 1e9 if km else None,
                              "call_GBps": b * 8 * shard * reps / (ms_b / 1e3) / 1e9}

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        mp = json.load(open(peaks_path))
        peak, peak_src = mp["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        bf16_peak, bf16_src = mp.get("bf16_tflops_sustained", mp["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bf16_peak, bf16_src = 1400.0, "fallback (B200_PROFILING.md, sustained)"
    for v in streaming.values():
        v["frac_of_hbm_peak"] = v["call_GBps"] / peak

    # BASELINE.json configs[1] as written (100 M codes, 1024-query batch, 1 B200), beside the metric's 1 B corpus
    config2 = None
    if world == 1 and n_total != 100_000_000:
        c2 = Corpus(ctx, _ffi.KIND_HAMMING64, 100_000_000)
        c2.append_synthetic(SEED_CORPUS, 0, 100_000_000)
        r2, k2 = planted(nq, 100_000_000)

        class _Rows2:
            __cuda_array_interface__ = {"shape": (100_000_000,), "typestr": "<i8", "data": (c2.device_rows_ptr(), False), "version": 2}

        v2 = torch.as_tensor(_Rows2(), device=dev)
        v2[torch.from_numpy(r2.astype(np.int64)).to(dev)] = torch.from_numpy(k2.view(np.int64)).to(dev)
        c2.refresh()
        for _ in range(3):
            c2.scan_hamming(q_dev, K, ids_out, dist_out)
        ms2 = timed(lambda: c2.scan_hamming(q_dev, K, ids_out, dist_out), 5) / 5
        config2 = {"workload": f"hamming top-{K}, {nq}-query batch over 100000000 synthetic 64-bit codes (BASELINE configs[1])",
                   "value": nq / (ms2 / 1e3), "unit": UNIT, "ms_per_step": ms2}
        del v2
        c2.close()

    # second half of BASELINE.json's metric: images hashed/s (multi bundle).  Every rank hashes its own batch
    # (the image batch simply splits across GPUs, no exchange); pixels resident in HBM, 408 B out per image.
    secondary = {}
    if not args.no_images:
        del view
        for (w, h, n_img) in ((1024, 1024, 1024), (256, 256, 8192)):
            px = torch.randint(0, 256, (n_img, h, w, 3), dtype=torch.uint8, device=dev)
            out = torch.zeros((n_img, 51), dtype=torch.int64, device=dev)
            for _ in range(3):
                ctx.image_hash_uniform(px, n_img, w, h, out=out)
            reps = 5
            ms_i = timed(lambda: ctx.image_hash_uniform(px, n_img, w, h, out=out), reps) / reps
            n_host = min(n_img, 128 if w == 1024 else 2048)
            px_host = px[:n_host].cpu().pin_memory()
            out_host = np.zeros((n_host, 51), dtype=np.uint64)
            for _ in range(2):
                ctx.image_hash_uniform(px_host.numpy(), n_host, w, h, out=out_host)
            ms_e = timed(lambda: ctx.image_hash_uniform(px_host.numpy(), n_host, w, h, out=out_host), 3) / 3
            img_parity = None
            if rank == 0 and args.parity_queries > 0:   # the batch that was just timed, first 32 images, against the oracle (docs/HASH_SPEC.md)
                import oracle
                want_words = oracle.image_multihash_batch(px[:32].cpu().numpy(), threads=oracle.host_threads())
                img_parity = {"images": 32, "ok": bool((out[:32].cpu().numpy().view(np.uint64) == want_words).all()),
                              "against": "oracle/ (docs/HASH_SPEC.md; parity with imgfprint 0.4.1 itself is unpinned)"}
            gbps = n_img * (3.0 * w * h + 408) / (ms_i / 1e3) / 1e9
            secondary[f"{w}x{h}"] = {"metric": "images_hashed_per_s_multi_bundle", "value": world * n_img / (ms_i / 1e3),
                                     "unit": "images/s", "images_per_gpu_per_step": n_img,
                                     "roofline": {"bound": "hbm", "achieved": gbps, "peak": peak, "unit": "GB/s", "frac": gbps / peak,
                                                  "note": "3*w*h + 408 algorithmic bytes per image; the spec's exact f32 arithmetic "
                                                          "(no FMA) makes the FP32/ALU issue rate the bound in force"},
                                     "e2e": {"value": world * n_host / (ms_e / 1e3), "unit": "images/s",
                                             "h2d_bytes_per_step": n_host * 3 * w * h, "d2h_bytes_per_step": n_host * 408},
                                     "parity_check": img_parity}
            del px, out, px_host
        try:
            secondary["jpeg_1024x1024"] = secondary_jpeg(ctx, rank, world, args.parity_queries > 0)
        except Exception as e:   # nvJPEG is loaded with dlopen at the first call: a box without it still benches the rest
            secondary["jpeg_1024x1024"] = {"unavailable": f"{type(e).__name__}: {e}"}
    # the other two scans of the hot path at BASELINE.json's shapes, sharded like the headline (configs[2], configs[3])
    if not args.no_paths:
        try:
            del view
        except NameError:
            pass
        corpus.close()
        torch.cuda.empty_cache()
        secondary["jaccard"] = secondary_jaccard(torch, ctx, group, rank, world, dev, timed, peak, args.small_paths)
        secondary["cosine"] = secondary_cosine(torch, ctx, group, rank, world, dev, timed, bf16_peak, args.small_paths)
    achieved = k_bytes / (k_ms / 1e3) / 1e9 if k_ms else None
    clk = clocks.summary()

    # Dominant kernel of the step.  A >= 64-query batch runs on the int8 tensor pipe: achieved = int8 operations the tensor
    # pipe EXECUTES (64 per (query, code) pair -- one 64-element +-1 dot product per TWO pairs) / kernel time.
    # Peak: MEASURED_PEAKS.json has no int8 figure, so the denominator is this repo's own measurement of the instruction the
    # kernel issues -- tcgen05.mma kind::i8, M128 x N256 x K64 back to back into two TMEM stages: 256.0 clk per tile
    # (scripts/micro/tmem_port.cu mode 0, output committed as profiles/r02_tmem_port.txt) = 16 384 int8 ops/clk/SM -- at the SM
    # clock sampled during this run.  The round-1 denominator (2 x the cuBLAS bf16 figure of MEASURED_PEAKS.json, taken at a
    # power-limited 1 305 MHz) is kept beside it as `frac_vs_2x_bf16_sustained`.
    pairs = nq * float(shard) * args.steps
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = clk["sm_mhz"] or 1965.0
    int8_peak = 16384.0 * sm_count * sm_mhz * 1e6 / 1e12
    if t_n:
        tops = t_ops / (t_ms / 1e3) / 1e12
        roofline = {"bound": "tensor", "kernel": "hamming_mma_scan_kernel", "achieved": tops, "peak": int8_peak,
                    "unit": "TFLOP/s", "frac": tops / int8_peak, "traffic": None,
                    "peak_source": f"measured: int8 tcgen05.mma M128xN256xK64 at 256.0 clk/tile (profiles/r02_tmem_port.txt) x {sm_count} SMs x "
                                   f"{sm_mhz:.0f} MHz sampled under load; MEASURED_PEAKS.json holds no int8 figure",
                    "frac_vs_2x_bf16_sustained": tops / (2 * bf16_peak), "tensor_pipe_active_ncu": committed_tensor_pipe_active(),
                    "launches": t_n, "kernel_ms_per_step": t_ms / args.steps,
                    "pairs_per_s": (t_ops / 64.0) / (t_ms / 1e3),
                    "all_scan_launches": {"launches": k_n, "kernel_ms_per_step": k_ms / args.steps,
                                          "pairs_per_s": pairs / (k_ms / 1e3) if k_ms else None},
                    "note": "achieved counts int8 operations (TOP/s) EXECUTED by the tensor pipe; `tensor_pipe_active_ncu` is "
                            "sm__pipe_tensor_cycles_active of the committed capture of this kernel (profiles/r02_hamming_mma_q1024_ncu.txt), "
                            "read from that file, not measured in this run.  The bound in force is the epilogue, not the MMAs: per 128 x 512-pair "
                            "accumulator item ~134 clk of tcgen05.ld plus ~430 clk of min/max work that do not overlap each other, around a "
                            "hand-over ring of ~380 clk (DESIGN.md 4.1, profiles/r02_hamming_schedules.md sections 5-6).  DRAM traffic is 8.02 B "
                            "per row per batch (ncu, profiles/hamming_scan_traffic.json): not measured live, hence `traffic` null.  `streaming` "
                            "is hamming_scan_kernel with 1-2 queries per corpus pass, where HBM is the bound",
                    "streaming": streaming}
    else:
        roofline = {"bound": "hbm", "kernel": "hamming_scan_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak if achieved else None, "traffic": None, "peak_source": peak_src, "launches": k_n,
                    "kernel_ms_per_step": k_ms / args.steps, "pairs_per_s": pairs / (k_ms / 1e3) if k_ms else None,
                    "note": "algorithmic bytes = 8 B x rows x queries per launch", "streaming": streaming}

    # Parity of the configuration that was measured, on the hardware path that was measured (N ranks, NCCL gather,
    # merge): the CPU oracle scans ALL n_total rows for every (nq / parity_queries)-th query and both the device-resident
    # and the end-to-end answers must equal it byte for byte.  Untimed; rank 0 only (every rank holds the same answer).
    parity = None
    if rank == 0 and args.parity_queries > 0:
        t0 = time.perf_counter()
        sel = np.arange(nq)[:: max(nq // args.parity_queries, 1)][: args.parity_queries]
        want_i, want_d = oracle_topk_full_corpus(n_total, nq, sel)
        ok_dev = bool((got_ids_dev[sel] == want_i).all() and (got_dist_dev[sel] == want_d).all())
        ok_e2e = bool((got_ids_e2e[sel] == want_i).all() and (got_dist_e2e[sel] == want_d).all())
        all_same = bool((got_ids_dev == got_ids_e2e).all() and (got_dist_dev == got_dist_e2e).all())
        parity = {"queries": int(len(sel)), "rows": n_total, "ok": ok_dev and ok_e2e and all_same, "device_path_ok": ok_dev,
                  "e2e_path_ok": ok_e2e, "device_equals_e2e_all_queries": all_same, "n_gpus": world,
                  "against": "oracle/ (CPU restatement of docs/HASH_SPEC.md section 6) over the full corpus, regenerated on the host",
                  "seconds": round(time.perf_counter() - t0, 1)}
        if not parity["ok"]:
            bad = [int(x) for x in sel[((got_ids_dev[sel] != want_i) | (got_dist_dev[sel] != want_d)).any(axis=1)]][:8]
            sys.stderr.write(f"PARITY FAILURE against the oracle: {parity}; first differing queries {bad}\n")

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_arm(n_total, nq, 1, 1, int(args.cpu_sample_rows), args.cpu_sample_queries)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": bench_config(n_total, nq, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nq * 8, "d2h_bytes_per_step": nq * K * 12,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(lt.item()),
            "roofline": roofline,
            "clocks": clk,
            "secondary": secondary,
            "config2": config2,
        }
        if cb is not None:
            line["cpu_baseline"] = cb
        if parity is not None:
            line["parity_check"] = parity
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    secondary_bad = [k for k, v in secondary.items() if isinstance(v, dict) and isinstance(v.get("parity_check"), dict) and not v["parity_check"]["ok"]]
    if (parity is not None and not parity["ok"]) or secondary_bad:
        sys.stderr.write(f"PARITY FAILURE: headline {parity}, secondary paths {secondary_bad}\n")
        return 3   # a fast answer that differs from the oracle is not a result
    return 0


if __name__ == "__main__":
    sys.exit(main())
