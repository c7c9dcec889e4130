import sys
import numpy as np
sys.path.insert(0, ".")
import oracle
from ucfp_b200 import Context, Corpus, _ffi
U64 = np.uint64
ctx = Context(0)
n, k = 1_300_000, 10
codes = oracle.fill_u64(n, 77)
codes[700_000:] = U64(0xFEEDFACECAFEBEEF)
ids = np.arange(n, 0, -1, dtype=U64)
queries = np.concatenate([np.array([0xFEEDFACECAFEBEEF, 0xFEEDFACECAFEBEEE], dtype=U64), oracle.fill_u64(30, 78)])
oi, od = oracle.hamming_topk(codes, queries, k, ids=ids, threads=oracle.host_threads())
for rep in range(40):
    c = Corpus(ctx, _ffi.KIND_HAMMING64, n); c.append(codes, ids)
    gi, gd = c.scan_hamming(queries, k)
    ok = bool((gi == oi).all() and (gd == od).all())
    ex = ctx.last_scan_exact_selects()
    print(rep, "fallbacks/fill", ctx.last_scan_stats(), "exact", ex & 0xFFFF, "longest re-scan list", (ex >> 16) & 0xFFFFFFFF, "rescan compactions", ex >> 48, "parity", ok, flush=True)
    c.close()
