"""Developer timing sweep for the Hamming scan (not the contract bench; see bench.py)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from ucfp_b200 import Context, Corpus, _ffi

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
qs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 16, 128, 1024]
ctx = Context(0)
corpus = Corpus(ctx, _ffi.KIND_HAMMING64, n)
corpus.append_synthetic(0xC0DE, 0, n)
torch.cuda.synchronize()
for nq in qs:
    q = torch.from_numpy(oracle.fill_u64(nq, 0xBEEF).view(np.int64)).cuda()
    ids = torch.empty((nq, 10), dtype=torch.int64, device="cuda"); d = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
    for _ in range(2): corpus.scan_hamming(q, 10, ids, d)
    torch.cuda.synchronize()
    reps = 5 if nq >= 128 else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): corpus.scan_hamming(q, 10, ids, d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"n": n, "nq": nq, "ms": round(ms, 4), "qps": round(nq / ms * 1e3, 1),
                      "alg_GBps": round(nq * 8 * n / ms / 1e6, 1), "Gpairs_s": round(nq * n / ms / 1e6, 1)}), flush=True)
