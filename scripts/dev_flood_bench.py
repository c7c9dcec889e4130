"""Developer timing of the overflow path: a Hamming corpus with a flood of identical codes whose record ids DESCEND with the row
(the adversarial order: every flood row beats the current k-th result, so the candidate lists of the queries that hit the flood
overflow in every chunk).  Run once as is and once with UCFP_RESCAN_ROUNDS=0 to compare the re-scan rounds with the exact
multi-pass selection.  usage: dev_flood_bench.py [rows] [flood_rows] [queries] [queries_hitting_the_flood]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from ucfp_b200 import Context, Corpus, _ffi

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
flood = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
hit = int(sys.argv[4]) if len(sys.argv) > 4 else 8
U64 = np.uint64
codes = oracle.fill_u64(n, 0xF100D)
code = U64(0xFEEDFACECAFEBEEF)
rng = np.random.default_rng(5)
rows = rng.choice(n, flood, replace=False)
codes[rows] = code
ids = np.arange(n, 0, -1, dtype=U64)           # descending with the row
queries = oracle.fill_u64(nq, 0xBEEF)
for j in range(hit):
    queries[j] = code ^ U64((1 << j) - 1)       # distance j to every flood row
ctx = Context(0)
corpus = Corpus(ctx, _ffi.KIND_HAMMING64, n)
corpus.append(codes, ids)
q = torch.from_numpy(queries.view(np.int64)).cuda()
oi = torch.empty((nq, 10), dtype=torch.int64, device="cuda"); od = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
for _ in range(2): corpus.scan_hamming(q, 10, oi, od)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps): corpus.scan_hamming(q, 10, oi, od)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fb, fill = ctx.last_scan_stats()
ex = ctx.last_scan_exact_selects()
sel = np.r_[0:hit, hit:hit + 8]
ri, rd = oracle.hamming_topk(codes, queries[sel], 10, ids=ids, threads=oracle.host_threads())
ok = bool((oi.cpu().numpy().view(U64)[sel] == ri).all() and (od.cpu().numpy().view(np.uint32)[sel] == rd).all())
print(json.dumps({"rows": n, "flood_rows": flood, "queries": nq, "queries_hitting_flood": hit, "rescan_rounds": os.environ.get("UCFP_RESCAN_ROUNDS", "default (2)"),
                  "ms_per_batch": round(ms, 3), "overflowed_queries": fb, "exact_selects": ex, "max_list_fill": fill, "parity_ok": ok}))
