"""One fixed scan for ncu (developer tool): python scripts/prof_scan.py cosine|jaccard N NQ"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from ucfp_b200 import Context, Corpus, _ffi
kind, n, nq = sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3])
ctx = Context(0)
if kind == "cosine":
    dim = 512
    corpus = Corpus(ctx, _ffi.KIND_COSINE, n, dim=dim)
    g = torch.Generator(device="cuda").manual_seed(1)
    for lo in range(0, n, 500_000):
        m = min(500_000, n - lo)
        x = torch.randn((m, dim), device="cuda", generator=g); x /= x.norm(dim=1, keepdim=True)
        corpus.append(x)
    q = torch.randn((nq, dim), device="cuda", generator=g)
    for _ in range(3): ids, sc = corpus.scan_cosine(q, 10)
else:
    corpus = Corpus(ctx, _ffi.KIND_MINHASH128, n)
    corpus.append_synthetic(0x5EED, 0, n)
    qh = oracle.fill_u64(nq * 128, 77).reshape(nq, 128)
    # plant: make a few corpus rows near-copies of each query so thresholds are realistic
    class _A:
        __cuda_array_interface__ = {"shape": (n, 128), "typestr": "<i8", "data": (corpus.device_rows_ptr(), False), "version": 2}
    view = torch.as_tensor(_A(), device="cuda")
    rng = np.random.default_rng(0)
    rows = rng.choice(n, n // 100, replace=False)
    base = oracle.fill_u64(len(rows) * 128, 99).reshape(-1, 128)
    qi = rng.integers(0, nq, len(rows)); p = rng.choice([0.9, 0.7, 0.5], len(rows))
    mask = rng.random((len(rows), 128)) < p[:, None]
    base[mask] = qh[qi][mask]
    view[torch.from_numpy(rows).cuda()] = torch.from_numpy(base.view(np.int64)).cuda()
    full = view.clone(); corpus.clear(); corpus.append(full); del full
    q = torch.from_numpy(qh.view(np.int64)).cuda()
    for _ in range(3): ids, m = corpus.scan_jaccard(q, 10)
torch.cuda.synchronize()
print("ok", ctx.last_scan_fallbacks())
