import sys, numpy as np
sys.path.insert(0, ".")
import oracle
from ucfp_b200 import Context, Corpus, _ffi
U64 = np.uint64
ctx = Context(0)
n, nq, kp, k = 300_000, 130, 64, 10
rng = np.random.default_rng(n)
rows = oracle.fill_u64(n * 51, 3).reshape(n, 51)
queries = oracle.fill_u64(nq * 51, 4).reshape(nq, 51)
def near(b, flips):
    out = b.copy()
    for w in range(51):
        for bit in rng.choice(64, size=rng.integers(0, flips + 1), replace=False):
            out[w] ^= U64(1) << U64(bit)
    return out
for j in range(nq):
    for f in (0, 2, 6, 12):
        rows[rng.integers(0, n)] = near(queries[j], f)
ids = rng.permutation(10 * n)[:n].astype(U64) + U64(7)
# coarse alone
hc = Corpus(ctx, _ffi.KIND_HAMMING64, n); hc.append(np.ascontiguousarray(rows[:, 17]), ids)
gi, gd = hc.scan_hamming(np.ascontiguousarray(queries[:, 17]), kp)
oi, od = oracle.hamming_topk(np.ascontiguousarray(rows[:, 17]), np.ascontiguousarray(queries[:, 17]), kp, ids=ids, threads=8)
print("coarse alone equal:", (gi == oi).all(), (gd == od).all(), "fallbacks", ctx.last_scan_stats())
mc = Corpus(ctx, _ffi.KIND_MULTIHASH, n); mc.append(rows, ids)
mi, ms = mc.scan_multihash(queries, kp, k)
wi, ws = oracle.multihash_rerank(rows, queries, kp, k, ids=ids, threads=8)
bad = np.where((mi != wi).any(axis=1) | (ms.view(np.uint32) != ws.view(np.uint32)).any(axis=1))[0]
print("rerank mismatching queries:", bad[:10], len(bad))
for q in bad[:3]:
    print(q, "got", mi[q], ms[q]); print(q, "want", wi[q], ws[q])
