import sys
import numpy as np
sys.path.insert(0, ".")
import oracle
from ucfp_b200 import Context, Corpus, _ffi
U64 = np.uint64
ctx = Context(0)
n = 300_000
for variant in (0, 1):
    codes = np.full(n, 0x0123456789ABCDEF, dtype=U64)
    if variant: codes[::2] ^= U64(1)
    k = 33 if variant else 10
    ids = np.arange(n, 0, -1, dtype=U64) * U64(3)
    queries = np.array([0x0123456789ABCDEF, 0x0123456789ABCDEE, 0], dtype=U64)
    oi, od = oracle.hamming_topk(codes, queries, k, ids=ids, threads=oracle.host_threads())
    c = Corpus(ctx, _ffi.KIND_HAMMING64, n); c.append(codes, ids)
    for rep in range(12):
        gi, gd = c.scan_hamming(queries, k)
        ok = bool((gi == oi).all() and (gd == od).all())
        print(variant, rep, "fallbacks/fill", ctx.last_scan_stats(), "exact", ctx.last_scan_exact_selects(), "parity", ok, flush=True)
    c.close()
