"""Behaviour of the C ABI beyond single calls: batches larger than one pass, incremental appends, clear/reuse,
large k, two contexts, concurrent callers.  Every result is still checked against the oracle."""
import threading

import numpy as np
import pytest

import oracle
from ucfp_b200 import Context, Corpus, UcfpError, _ffi

pytestmark = pytest.mark.gpu
U64 = np.uint64


def test_hamming_more_queries_than_one_pass_and_large_k(ctx):
    n = 60_000
    codes = oracle.fill_u64(n, 31)
    q = oracle.fill_u64(2500, 32)                       # > 2048 queries: two corpus passes inside one call
    c = Corpus(ctx, _ffi.KIND_HAMMING64, n)
    c.append(codes)
    gi, gd = c.scan_hamming(q, 10)
    oi, od = oracle.hamming_topk(codes, q, 10, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gd, od)
    gi, gd = c.scan_hamming(q[:3].copy(), 2048)         # the largest supported k
    oi, od = oracle.hamming_topk(codes, q[:3], 2048, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gd, od)
    with pytest.raises(UcfpError) as e:
        c.scan_hamming(q[:1].copy(), 2049)
    assert e.value.code == _ffi.E_UNSUPPORTED
    c.close()


def test_jaccard_and_cosine_multi_pass_batches(ctx):
    sig = oracle.fill_u64(8000 * 128, 3).reshape(-1, 128)
    q = oracle.fill_u64(300 * 128, 4).reshape(300, 128)  # > 256 queries per pass
    for j in range(300):
        sig[(j * 17) % 8000, : 30 + j % 90] = q[j, : 30 + j % 90]
    c = Corpus(ctx, _ffi.KIND_MINHASH128, len(sig))
    c.append(sig)
    gi, gm = c.scan_jaccard(q, 5)
    oi, om = oracle.jaccard_topk(sig, q, 5, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gm, om)
    c.close()
    rng = np.random.default_rng(1)
    rows = rng.standard_normal((20_000, 96)).astype(np.float32)
    qv = rng.standard_normal((1100, 96)).astype(np.float32)  # > 1024 queries per pass
    c = Corpus(ctx, _ffi.KIND_COSINE, len(rows), dim=96)
    c.append(rows)
    gi, gs = c.scan_cosine(qv, 10)
    oi, osc, _ = oracle.cosine_topk(rows, qv, 10, mode=1, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gs.view(np.uint32), osc.view(np.uint32))
    c.close()


def test_incremental_append_clear_and_reuse(ctx):
    codes = oracle.fill_u64(50_000, 9)
    ids = (np.arange(50_000, dtype=U64) * U64(5)) + U64(1)
    q = oracle.fill_u64(7, 10)
    c = Corpus(ctx, _ffi.KIND_HAMMING64, 50_000)
    for lo, hi in ((0, 1), (1, 2049), (2049, 30_000), (30_000, 50_000)):     # upserts arrive in odd-sized batches
        c.append(codes[lo:hi].copy(), ids[lo:hi].copy())
        gi, gd = c.scan_hamming(q, 10)
        oi, od = oracle.hamming_topk(codes[:hi], q, 10, ids=ids[:hi])
        np.testing.assert_array_equal(gi, oi)
        np.testing.assert_array_equal(gd, od)
    assert len(c) == 50_000
    c.clear()
    assert len(c) == 0
    c.append(codes[:100].copy())                                               # implicit ids after a clear
    gi, _ = c.scan_hamming(q, 3)
    oi, _ = oracle.hamming_topk(codes[:100], q, 3)
    np.testing.assert_array_equal(gi, oi)
    c.close()
    rng = np.random.default_rng(2)
    rows = rng.standard_normal((3000, 64)).astype(np.float32)
    c = Corpus(ctx, _ffi.KIND_COSINE, 3000, dim=64)
    for lo in range(0, 3000, 701):
        c.append(rows[lo:lo + 701].copy())
    gi, gs = c.scan_cosine(rows[:4].copy(), 5)
    oi, osc, _ = oracle.cosine_topk(rows, rows[:4], 5, mode=1)
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gs.view(np.uint32), osc.view(np.uint32))
    assert (gi[:, 0] == np.arange(4)).all()                                    # every row is its own best match
    c.close()


def test_two_contexts_and_concurrent_callers(ctx):
    """Entry points are thread-safe (the reference calls them from up to 512 tokio tasks, bin/ucfp.rs:267)."""
    other = Context(0, use_torch_stream=False)           # its own non-blocking stream
    codes = oracle.fill_u64(200_000, 77)
    q = oracle.fill_u64(64, 78)
    want = oracle.hamming_topk(codes, q, 10, threads=oracle.host_threads())
    corpora = [Corpus(ctx, _ffi.KIND_HAMMING64, len(codes)), Corpus(other, _ffi.KIND_HAMMING64, len(codes))]
    for c in corpora:
        c.append(codes)
    errors = []

    def worker(c, lo):
        try:
            for _ in range(5):
                gi, gd = c.scan_hamming(q[lo:lo + 16].copy(), 10)
                np.testing.assert_array_equal(gi, want[0][lo:lo + 16])
                np.testing.assert_array_equal(gd, want[1][lo:lo + 16])
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(corpora[i % 2], 16 * i)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for c in corpora:
        c.close()
    other.close()


def test_invalid_arguments_are_errors_not_crashes(ctx):
    with pytest.raises(UcfpError) as e:
        Corpus(ctx, 99, 10)
    assert e.value.code == _ffi.E_INVALID
    with pytest.raises(UcfpError):
        Corpus(ctx, _ffi.KIND_COSINE, 10, dim=0)
    with pytest.raises(UcfpError):
        Corpus(ctx, _ffi.KIND_HAMMING64, 0)
    c = Corpus(ctx, _ffi.KIND_COSINE, 10, dim=8)
    with pytest.raises(UcfpError) as e:
        c.scan_hamming(np.zeros(1, U64), 1)
    assert e.value.code == _ffi.E_STATE
    c.close()
    got, status = ctx.image_hash_batch([np.zeros((3, 100, 3), np.uint8)])      # height below the 4x4 block grid
    assert status[0] == _ffi.E_INVALID and (got == 0).all()
