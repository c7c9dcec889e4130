"""Developer timing for the Jaccard scan (not the contract bench)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from ucfp_b200 import Context, Corpus, _ffi
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
ctx = Context(0)
corpus = Corpus(ctx, _ffi.KIND_MINHASH128, n)
corpus.append_synthetic(0x5EED, 0, n)
torch.cuda.synchronize()
planted = "--plant" in sys.argv
qall = oracle.fill_u64(256 * 128, 77).reshape(256, 128)
if planted:  # BASELINE config 3: 1 % of the rows copy a random query's slots with p in {.9,.7,.5}
    class _A:
        __cuda_array_interface__ = {"shape": (n, 128), "typestr": "<i8", "data": (corpus.device_rows_ptr(), False), "version": 2}
    view = torch.as_tensor(_A(), device="cuda")
    rng = np.random.default_rng(0)
    rows = rng.choice(n, n // 100, replace=False)
    base = oracle.fill_u64(len(rows) * 128, 99).reshape(-1, 128)
    qi = rng.integers(0, 256, len(rows)); p = rng.choice([0.9, 0.7, 0.5], len(rows))
    mask = rng.random((len(rows), 128)) < p[:, None]
    base[mask] = qall[qi][mask]
    view[torch.from_numpy(rows).cuda()] = torch.from_numpy(base.view(np.int64)).cuda()
    # the sketch of the overwritten rows must be rebuilt: re-append is not possible, so use a fresh corpus
    full = view.clone(); corpus.clear(); corpus.append(full); del full
for nq in (1, 8, 64, 256):
    qh = qall[:nq].copy()
    q = torch.from_numpy(qh.view(np.int64)).cuda()
    # plant neighbours of each query into the corpus? (synthetic rows only here: the no-neighbour worst case)
    ids = torch.empty((nq, 10), dtype=torch.int64, device="cuda"); m = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
    for _ in range(2): corpus.scan_jaccard(q, 10, ids, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps): corpus.scan_jaccard(q, 10, ids, m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"n": n, "nq": nq, "ms": round(ms, 3), "qps": round(nq / ms * 1e3, 1), "alg_GBps": round(nq * 1024 * n / ms / 1e6, 1),
                      "Gpairs_s": round(nq * n / ms / 1e6, 2)}), flush=True)
