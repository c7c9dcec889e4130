import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from ucfp_b200 import Context, Corpus, _ffi
n, nq = int(float(sys.argv[1])), int(sys.argv[2])
ctx = Context(0)
corpus = Corpus(ctx, _ffi.KIND_MINHASH128, n)
corpus.append_synthetic(0x5EED, 0, n)
qh = oracle.fill_u64(nq * 128, 77).reshape(nq, 128)
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
corpus.refresh()
q = torch.from_numpy(qh.view(np.int64)).cuda()
for _ in range(3): ids, m = corpus.scan_jaccard(q, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): corpus.scan_jaccard(q, 10)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"n": n, "nq": nq, "ms": round(ms, 3), "qps": round(nq / ms * 1e3, 1), "fallbacks": ctx.last_scan_fallbacks()}))
