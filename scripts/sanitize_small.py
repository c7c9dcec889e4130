"""One tiny invocation of every kernel family, for compute-sanitizer memcheck (developer tool)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import oracle
from ucfp_b200 import Context, Corpus, _ffi
ctx = Context(0)
codes = oracle.fill_u64(70_001, 1); q = oracle.fill_u64(5, 2)
c = Corpus(ctx, _ffi.KIND_HAMMING64, len(codes)); c.append(codes)
i, d = c.scan_hamming(q, 10); oi, od = oracle.hamming_topk(codes, q, 10); assert (i == oi).all() and (d == od).all(); c.close()
flood = np.full(9000, 7, np.uint64); ids = np.arange(9000, 0, -1).astype(np.uint64)
c = Corpus(ctx, _ffi.KIND_HAMMING64, 9000); c.append(flood, ids); i, d = c.scan_hamming(np.array([7], np.uint64), 3); assert i[0].tolist() == [1, 2, 3]; c.close()
sig = oracle.fill_u64(3001 * 128, 3).reshape(-1, 128); qs = oracle.fill_u64(2 * 128, 4).reshape(2, 128); sig[17, :80] = qs[0, :80]
c = Corpus(ctx, _ffi.KIND_MINHASH128, len(sig)); c.append(sig)
i, m = c.scan_jaccard(qs, 5); oi, om = oracle.jaccard_topk(sig, qs, 5); assert (i == oi).all() and (m == om).all(); c.close()
rng = np.random.default_rng(0); rows = rng.standard_normal((5000, 100)).astype(np.float32); qv = rng.standard_normal((3, 100)).astype(np.float32)
c = Corpus(ctx, _ffi.KIND_COSINE, len(rows), dim=100); c.append(rows)
i, s = c.scan_cosine(qv, 10); oi, os_, _ = oracle.cosine_topk(rows, qv, 10, mode=1); assert (i == oi).all() and (s.view(np.uint32) == os_.view(np.uint32)).all(); c.close()
imgs = [oracle.fill_u64(256 * 256 * 3 // 8, 5).view(np.uint8).reshape(256, 256, 3).copy(), oracle.fill_u64(1024 * 160 * 3 // 8, 6).view(np.uint8).reshape(160, 1024, 3).copy(),
        oracle.fill_u64((37 * 53 * 3 + 7) // 8, 7).view(np.uint8)[: 37 * 53 * 3].reshape(53, 37, 3).copy()]
got, st = ctx.image_hash_batch(imgs); assert (st == 0).all()
for g, im in zip(got, imgs): assert (g == oracle.image_multihash(im)).all()
mi, mk = np.zeros((1, 2), np.uint64), np.zeros((1, 2), np.uint32)
ctx.merge_topk_u32(np.array([[[3, 9]], [[1, 5]]], np.uint64), np.array([[[1, 4]], [[1, 2]]], np.uint32), 2, 1, 2, False, mi, mk); assert mi[0].tolist() == [1, 3]
# round 2: tensor scan with parked strips + recheck, histogram compaction, two-columns-per-thread image kernel, mutation, multi-hash
codes = oracle.fill_u64(150_000, 11); q = oracle.fill_u64(130, 12)
for j in range(20): codes[1000 * j + 17] = q[j] ^ np.uint64(5)
c = Corpus(ctx, _ffi.KIND_HAMMING64, len(codes)); c.append(codes)
i, d = c.scan_hamming(q, 10); oi, od = oracle.hamming_topk(codes, q, 10, threads=4); assert (i == oi).all() and (d == od).all()
ids = np.arange(len(codes), dtype=np.uint64)
assert c.delete(ids[100:164]) == 64
assert c.upsert(np.array([5, 10**7], np.uint64), np.array([q[0], q[1]], np.uint64)) == 1
i, d = c.scan_hamming(q[:70], 3); assert d[0, 0] == 0 and d[1, 0] == 0; c.close()
imgs = [oracle.fill_u64((130 * 200 * 3 + 7) // 8, 8).view(np.uint8)[: 130 * 200 * 3].reshape(200, 130, 3).copy(),
        oracle.fill_u64(384 * 216 * 3 // 8, 9).view(np.uint8).reshape(216, 384, 3).copy()]
got, st = ctx.image_hash_batch(imgs); assert (st == 0).all()
for g, im in zip(got, imgs): assert (g == oracle.image_multihash(im)).all()
b = oracle.fill_u64(3000 * 51, 13).reshape(-1, 51); qb = oracle.fill_u64(2 * 51, 14).reshape(-1, 51); b[77] = qb[0]
c = Corpus(ctx, _ffi.KIND_MULTIHASH, len(b)); c.append(b)
gi, gs = c.scan_multihash(qb, 16, 4); oi, osc = oracle.multihash_rerank(b, qb, 16, 4); assert (gi == oi).all() and (gs.view(np.uint32) == osc.view(np.uint32)).all(); c.close()
print("sanitize_small ok")
