"""Developer timing for the cosine scan (not the contract bench)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from ucfp_b200 import Context, Corpus, _ffi
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
dim = 512
ctx = Context(0)
corpus = Corpus(ctx, _ffi.KIND_COSINE, n, dim=dim)
g = torch.Generator(device="cuda").manual_seed(1)
for lo in range(0, n, 500_000):
    m = min(500_000, n - lo)
    x = torch.randn((m, dim), device="cuda", generator=g); x /= x.norm(dim=1, keepdim=True)
    corpus.append(x)
torch.cuda.synchronize()
for nq in (1, 16, 128, 256, 1024):
    q = torch.randn((nq, dim), device="cuda", generator=g); q /= q.norm(dim=1, keepdim=True)
    ids = torch.empty((nq, 10), dtype=torch.int64, device="cuda"); sc = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
    for _ in range(2): corpus.scan_cosine(q, 10, ids, sc)
    fb = ctx.last_scan_fallbacks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    ctx.profile_begin()
    e0.record()
    for _ in range(reps): corpus.scan_cosine(q, 10, ids, sc)
    e1.record(); torch.cuda.synchronize()
    kms, kfl, kn = ctx.profile_end(_ffi.PROF_COSINE_SCAN)
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"n": n, "nq": nq, "ms": round(ms, 3), "qps": round(nq / ms * 1e3, 1), "TFLOPs_call": round(2 * n * dim * nq / ms / 1e9, 1),
                      "TFLOPs_kernel": round(kfl / kms / 1e9, 1), "fallbacks": fb}), flush=True)
