"""BASELINE.json configs 3, 4 and 5 at sizes a test can afford (configs 1 and 2 run at full size in
test_image_gpu.py / test_hamming_gpu.py; the full 50 M / 20 M / 8 192-image shapes are run, with a parity check on
a sample, by scripts/bench_paths.py).  Each case combines size-independent properties over the whole result with
bit-exact agreement with the oracle on a window of the same corpus."""
import numpy as np
import pytest

import oracle
from ucfp_b200 import Corpus, _ffi

pytestmark = pytest.mark.gpu
U64 = np.uint64


def _view(corpus, shape, typestr):
    import torch

    class _A:
        __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (corpus.device_rows_ptr(), False), "version": 2}
    return torch.as_tensor(_A(), device="cuda")


def test_config3_jaccard_5m_signatures_256_queries(ctx):
    import torch
    n, nq, k = 5_000_000, 256, 10
    corpus = Corpus(ctx, _ffi.KIND_MINHASH128, n)
    corpus.append_synthetic(0x5EED, 0, n)
    q = oracle.fill_u64(nq * 128, 77).reshape(nq, 128)
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(n, n // 100, replace=False))
    base = oracle.fill_u64(len(rows) * 128, 99).reshape(-1, 128)
    qi, p = rng.integers(0, nq, len(rows)), rng.choice([0.9, 0.7, 0.5], len(rows))
    mask = rng.random((len(rows), 128)) < p[:, None]
    base[mask] = q[qi][mask]
    view = _view(corpus, (n, 128), "<i8")
    view[torch.from_numpy(rows).cuda()] = torch.from_numpy(base.view(np.int64)).cuda()
    corpus.refresh()                                    # sketches of the planted rows
    ids, m = corpus.scan_jaccard(torch.from_numpy(q.view(np.int64)).cuda(), k)
    torch.cuda.synchronize()
    ids, m = ids.cpu().numpy().view(U64), m.cpu().numpy().view(np.uint32)
    assert ctx.last_scan_fallbacks() == 0
    # ordered by (matches desc, id asc); every reported count is the true count of that row
    assert (np.diff(m.astype(np.int64), axis=1) <= 0).all()
    same = np.diff(m.astype(np.int64), axis=1) == 0
    assert (np.diff(ids.astype(np.int64), axis=1)[same] > 0).all()
    got_rows = view[torch.from_numpy(ids.astype(np.int64).ravel()).cuda()].cpu().numpy().view(U64).reshape(nq, k, 128)
    np.testing.assert_array_equal((got_rows == q[:, None, :]).sum(axis=2).astype(np.uint32), m)
    # no planted row with more matches than the k-th result is missing
    planted_m = (base == q[qi]).sum(axis=1)
    for j in range(0, nq, 5):
        better = rows[(qi == j) & (planted_m > m[j, -1])]
        assert np.isin(better, ids[j]).all(), j
    # bit-exact with the oracle on a window (first 300 k rows contain ~3 000 planted rows)
    win = view[:300_000].cpu().numpy().view(U64)
    sub = Corpus(ctx, _ffi.KIND_MINHASH128, len(win))
    sub.append(win)
    gi, gm = sub.scan_jaccard(q[:64].copy(), k)
    oi, om = oracle.jaccard_topk(win, q[:64], k, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gm, om)
    sub.close()
    corpus.close()


def test_config4_cosine_2m_vectors_1024_queries(ctx):
    import torch
    n, dim, nq, k = 2_000_000, 512, 1024, 10
    corpus = Corpus(ctx, _ffi.KIND_COSINE, n, dim=dim)
    g = torch.Generator(device="cuda").manual_seed(1)
    for lo in range(0, n, 500_000):                      # unit-norm rows with bf16-representable values (config 4)
        x = torch.randn((500_000, dim), device="cuda", generator=g)
        corpus.append((x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32))
    q = torch.randn((nq, dim), device="cuda", generator=g)
    q = (q / q.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32)
    view = _view(corpus, (n, dim), "<f4")
    prow = torch.randperm(n, device="cuda", generator=g)[: nq * 8]
    pv = q.repeat_interleave(8, dim=0) + 0.03 * torch.randn((nq * 8, dim), device="cuda", generator=g)
    view[prow] = (pv / pv.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32)
    corpus.refresh()
    ids, sc = corpus.scan_cosine(q, k)
    torch.cuda.synchronize()
    assert ctx.last_scan_fallbacks() == 0               # the tcgen05 coarse pass, not the fallback, selected
    ids_h, sc_h = ids.cpu().numpy().view(U64), sc.cpu().numpy()
    assert (np.diff(sc_h, axis=1) <= 0).all()
    # the 8 planted neighbours of every query (cosine ~ 0.8) beat all random rows (|cosine| < 0.3)
    planted = prow.cpu().numpy().astype(U64).reshape(nq, 8)
    assert all(np.isin(planted[j], ids_h[j]).all() for j in range(nq)), "planted neighbours missing"
    assert (sc_h[:, 7] > 0.6).all() and (sc_h[:, 8] < 0.4).all()
    # reported scores are the reference's f32 arithmetic, bit for bit (spot-check 64 queries against the oracle)
    sel = np.arange(0, nq, 16)
    rows_h = view[torch.from_numpy(ids_h[sel].astype(np.int64).ravel()).cuda()].cpu().numpy().reshape(len(sel), k, dim)
    qh = q.cpu().numpy()
    for a, j in enumerate(sel):
        for b in range(k):
            want = np.float32(oracle.dot_product(qh[j], rows_h[a, b])) / np.float32(np.float32(oracle.l2_norm(qh[j])) * np.float32(oracle.l2_norm(rows_h[a, b])))
            assert sc_h[j, b] == want
    # bit-exact ids and scores against the oracle on a 100 k-row window
    win = view[:100_000].cpu().numpy()
    sub = Corpus(ctx, _ffi.KIND_COSINE, len(win), dim=dim)
    sub.append(win)
    gi, gs = sub.scan_cosine(qh[:32].copy(), k)
    oi, osc, _ = oracle.cosine_topk(win, qh[:32], k, mode=1, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gs.view(np.uint32), osc.view(np.uint32))
    sub.close()
    corpus.close()


def test_config5_ingest_hashing_1024x1024_images(ctx):
    """A chunk of config 5: synthetic 1024 x 1024 RGB images, multi bundle, device-resident batch."""
    import torch
    n, w, h = 96, 1024, 1024
    host = oracle.fill_u64(n * w * h * 3 // 8, 0x1316).view(np.uint8).reshape(n, h, w, 3).copy()
    y, x = np.mgrid[0:h, 0:w]
    host[0] = np.stack([x % 256, y % 256, np.full_like(x, 128)], -1)           # the reference ramp
    host[1] = np.clip(host[1].astype(np.int32) // 4 + (x[..., None] // 8), 0, 255).astype(np.uint8)   # smooth gradient + noise
    out = ctx.image_hash_uniform(torch.from_numpy(host).cuda(), n, w, h).cpu().numpy().view(U64)
    want = oracle.image_multihash_batch(host, threads=oracle.host_threads())
    np.testing.assert_array_equal(out, want)
    out_h = ctx.image_hash_uniform(host, n, w, h)                               # the same batch from host memory
    np.testing.assert_array_equal(out_h, want)


def test_widest_supported_images_take_the_generic_kernel(ctx):
    """max_dimension of the reference's PreprocessConfig is 8192 (algorithms_manifest.rs:446-469)."""
    img = oracle.fill_u64(8192 * 96 * 3 // 8, 5).view(np.uint8).reshape(96, 8192, 3).copy()
    tall = oracle.fill_u64(64 * 8192 * 3 // 8, 6).view(np.uint8).reshape(8192, 64, 3).copy()
    got, status = ctx.image_hash_batch([img, tall])
    assert (status == 0).all()
    np.testing.assert_array_equal(got[0], oracle.image_multihash(img))
    np.testing.assert_array_equal(got[1], oracle.image_multihash(tall))


# ---- configs 3 and 4 at FULL size, the oracle over ALL rows (the metric configurations themselves were never
# parity-checked).  The corpus is replayed on the host chunk by chunk, so host memory stays bounded; 16 of the queries are
# checked end to end (the oracle is O(rows x queries) on CPU cores), all of them through the size-independent properties.
def _merge_topk(best_i, best_k, new_i, new_k, k, descending):
    ci, ck = np.concatenate([best_i, new_i], axis=1), np.concatenate([best_k, new_k], axis=1)
    for r in range(len(ci)):
        valid = ci[r] != U64(2**64 - 1)
        primary = np.where(valid, -ck[r].astype(np.float64) if descending else ck[r].astype(np.float64), np.inf)
        order = np.lexsort((ci[r], primary))[:k]
        best_i[r], best_k[r] = ci[r][order], ck[r][order]


def test_config3_full_50m_signatures_oracle_over_all_rows(ctx):
    import torch
    n, nq, k, chunk = 50_000_000, 256, 10, 2_000_000
    corpus = Corpus(ctx, _ffi.KIND_MINHASH128, n)
    corpus.append_synthetic(0x5EED, 0, n)
    q = oracle.fill_u64(nq * 128, 77).reshape(nq, 128)
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(n, n // 100, replace=False))          # config 3: 1 % of the rows copy a query's slots with p in {.9,.7,.5}
    base = oracle.fill_u64(len(rows) * 128, 99).reshape(-1, 128)
    qi, p = rng.integers(0, nq, len(rows)), rng.choice([0.9, 0.7, 0.5], len(rows))
    for a in range(0, len(rows), 100_000):
        mask = rng.random((min(100_000, len(rows) - a), 128)) < p[a:a + 100_000, None]
        blk = base[a:a + 100_000]
        blk[mask] = q[qi[a:a + 100_000]][mask]
    view = _view(corpus, (n, 128), "<i8")
    for a in range(0, len(rows), 100_000):
        view[torch.from_numpy(rows[a:a + 100_000]).cuda()] = torch.from_numpy(base[a:a + 100_000].view(np.int64)).cuda()
    corpus.refresh()
    ids, m = corpus.scan_jaccard(torch.from_numpy(q.view(np.int64)).cuda(), k)
    torch.cuda.synchronize()
    ids, m = ids.cpu().numpy().view(U64), m.cpu().numpy().view(np.uint32)
    assert ctx.last_scan_fallbacks() == 0
    assert (np.diff(m.astype(np.int64), axis=1) <= 0).all()
    # the oracle over all 50 M rows for every 16th query
    sel = np.arange(0, nq, 16)
    qs = np.ascontiguousarray(q[sel])
    best_i = np.full((len(sel), k), U64(2**64 - 1), dtype=U64)
    best_m = np.zeros((len(sel), k), dtype=np.uint32)
    buf = np.empty(chunk * 128, dtype=U64)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        host = oracle.fill_u64((hi - lo) * 128, 0x5EED, start=lo * 128, out=buf).reshape(hi - lo, 128)
        a, b = np.searchsorted(rows, lo), np.searchsorted(rows, hi)
        host[rows[a:b] - lo] = base[a:b]
        oi, om = oracle.jaccard_topk(host, qs, k, id_base=lo, threads=oracle.host_threads())
        _merge_topk(best_i, best_m, oi, om, k, descending=True)
    np.testing.assert_array_equal(m[sel], best_m)
    np.testing.assert_array_equal(ids[sel], best_i)
    corpus.close()


def test_config4_full_20m_vectors_oracle_over_all_rows(ctx):
    import torch
    n, dim, nq, k, blk = 20_000_000, 512, 1024, 10, 500_000
    corpus = Corpus(ctx, _ffi.KIND_COSINE, n, dim=dim)
    g = torch.Generator(device="cuda").manual_seed(1)
    for lo in range(0, n, blk):                                      # unit-norm rows with bf16-representable values (config 4)
        x = torch.randn((blk, dim), device="cuda", generator=g)
        corpus.append((x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32))
    q = torch.randn((nq, dim), device="cuda", generator=g)
    q = (q / q.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32)
    view = _view(corpus, (n, dim), "<f4")
    prow = torch.randperm(n, device="cuda", generator=g)[: nq * 8]
    pv = q.repeat_interleave(8, dim=0) + 0.03 * torch.randn((nq * 8, dim), device="cuda", generator=g)
    view[prow] = (pv / pv.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(torch.float32)
    corpus.refresh()
    ids, sc = corpus.scan_cosine(q, k)
    torch.cuda.synchronize()
    assert ctx.last_scan_fallbacks() == 0
    ids_h, sc_h = ids.cpu().numpy().view(U64), sc.cpu().numpy()
    planted = prow.cpu().numpy().astype(U64).reshape(nq, 8)
    assert all(np.isin(planted[j], ids_h[j]).all() for j in range(nq)), "planted neighbours missing"
    # the oracle (restatement of src/index/embedded/mod.rs:268-360) over all 20 M rows for every 64th query: bit-exact f32 scores
    sel = np.arange(0, nq, 64)
    qs = np.ascontiguousarray(q.cpu().numpy()[sel])
    best_i = np.full((len(sel), k), U64(2**64 - 1), dtype=U64)
    best_s = np.full((len(sel), k), -np.inf, dtype=np.float32)
    for lo in range(0, n, 2 * blk):
        hi = min(n, lo + 2 * blk)
        host = view[lo:hi].cpu().numpy()
        oi, osc, _ = oracle.cosine_topk(host, qs, k, id_base=lo, mode=1, threads=oracle.host_threads())
        _merge_topk(best_i, best_s, oi, osc, k, descending=True)
    np.testing.assert_array_equal(ids_h[sel], best_i)
    np.testing.assert_array_equal(sc_h[sel].view(np.uint32), best_s.view(np.uint32))
    corpus.close()
