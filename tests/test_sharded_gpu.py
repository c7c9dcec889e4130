"""Multi-GPU protocol on one GPU: a corpus split into record-range shards (separate HBM corpora with
global ids), each scanned on its own, per-shard top-k merged by ucfp_merge_topk_* -- the step that follows the
NCCL all-gather -- must be byte-identical to scanning the unsplit corpus, for all three scans."""
import numpy as np
import pytest

import oracle
from ucfp_b200 import Corpus, _ffi
from ucfp_b200.sharding import merge_topk_host, shard_range

pytestmark = pytest.mark.gpu
U64 = np.uint64


def _sharded(ctx, kind, rows, queries, k, parts, dim=0, explicit_ids=None):
    n = len(rows)
    ids_parts, key_parts = [], []
    for r in range(parts):
        lo, hi = shard_range(n, r, parts)
        c = Corpus(ctx, kind, max(hi - lo, 1), dim=dim)
        if explicit_ids is None:
            c.set_id_base(lo)
            if hi > lo:
                c.append(np.ascontiguousarray(rows[lo:hi]))
        elif hi > lo:
            c.append(np.ascontiguousarray(rows[lo:hi]), np.ascontiguousarray(explicit_ids[lo:hi]))
        scan = {_ffi.KIND_HAMMING64: c.scan_hamming, _ffi.KIND_MINHASH128: c.scan_jaccard, _ffi.KIND_COSINE: c.scan_cosine}[kind]
        i, v = scan(queries, k)
        ids_parts.append(i)
        key_parts.append(v)
        c.close()
    return np.stack(ids_parts), np.stack(key_parts)


@pytest.mark.parametrize("parts", [2, 4, 8])
def test_hamming_shards_merge_to_single_corpus_result(ctx, parts):
    n, nq, k = 200_003, 33, 10
    codes = oracle.fill_u64(n, 5) & U64(0x3FFFFF)          # few distinct bits: ties straddle shard boundaries
    q = oracle.fill_u64(nq, 6) & U64(0x3FFFFF)
    ids_all, d_all = _sharded(ctx, _ffi.KIND_HAMMING64, codes, q, k, parts)
    mi, md = np.zeros((nq, k), U64), np.zeros((nq, k), np.uint32)
    ctx.merge_topk_u32(ids_all, d_all, parts, nq, k, False, mi, md)
    oi, od = oracle.hamming_topk(codes, q, k, threads=oracle.host_threads())
    np.testing.assert_array_equal(mi, oi)
    np.testing.assert_array_equal(md, od)
    hi_, hd_ = merge_topk_host(ids_all, d_all, k, descending=False)   # the CPU statement of the same merge (gloo test)
    np.testing.assert_array_equal(mi, hi_)
    np.testing.assert_array_equal(md, hd_)


def test_jaccard_shards_with_explicit_ids(ctx):
    n, nq, k, parts = 40_000, 6, 10, 4
    sig = oracle.fill_u64(n * 128, 7).reshape(n, 128)
    q = oracle.fill_u64(nq * 128, 8).reshape(nq, 128)
    rng = np.random.default_rng(2)
    for j in range(nq):
        for r in rng.choice(n, 30, replace=False):
            m = rng.random(128) < 0.6
            sig[r, m] = q[j, m]
    ids = rng.permutation(n).astype(U64) * U64(7)
    ids_all, m_all = _sharded(ctx, _ffi.KIND_MINHASH128, sig, q, k, parts, explicit_ids=ids)
    mi, mm = np.zeros((nq, k), U64), np.zeros((nq, k), np.uint32)
    ctx.merge_topk_u32(ids_all, m_all, parts, nq, k, True, mi, mm)
    oi, om = oracle.jaccard_topk(sig, q, k, ids=ids, threads=oracle.host_threads())
    np.testing.assert_array_equal(mi, oi)
    np.testing.assert_array_equal(mm, om)


def test_cosine_shards_merge_f32(ctx):
    n, dim, nq, k, parts = 60_000, 128, 20, 10, 3
    rng = np.random.default_rng(4)
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    rows[1000:1200] = rows[:200]                           # exact duplicates: equal scores across shards
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    ids_all, s_all = _sharded(ctx, _ffi.KIND_COSINE, rows, q, k, parts, dim=dim)
    mi, ms = np.zeros((nq, k), U64), np.zeros((nq, k), np.float32)
    ctx.merge_topk_f32(ids_all, s_all, parts, nq, k, mi, ms)
    oi, osc, _ = oracle.cosine_topk(rows, q, k, mode=1, threads=oracle.host_threads())
    np.testing.assert_array_equal(mi, oi)
    np.testing.assert_array_equal(ms.view(np.uint32), osc.view(np.uint32))


def test_merge_keeps_sentinels_last_and_handles_short_lists(ctx):
    """Shards with fewer than k rows pad with (UINT64_MAX, sentinel); the merge must drop those first."""
    parts, nq, k = 3, 2, 4
    NONE, S32 = U64(2**64 - 1), np.uint32(2**32 - 1)
    ids = np.full((parts, nq, k), NONE, U64)
    keys = np.full((parts, nq, k), S32, np.uint32)
    ids[0, 0, :2], keys[0, 0, :2] = [5, 9], [1, 3]
    ids[2, 0, :1], keys[2, 0, :1] = [7], [1]
    ids[1, 1, :3], keys[1, 1, :3] = [4, 2, 8], [0, 6, 6]
    mi, mk = np.zeros((nq, k), U64), np.zeros((nq, k), np.uint32)
    ctx.merge_topk_u32(ids, keys, parts, nq, k, False, mi, mk)
    assert mi[0].tolist() == [5, 7, 9, int(NONE)] and mk[0].tolist() == [1, 1, 3, int(S32)]
    assert mi[1].tolist() == [4, 2, 8, int(NONE)] and mk[1].tolist() == [0, 6, 6, int(S32)]
    ctx.merge_topk_u32(ids, keys, parts, nq, k, True, mi, mk)          # descending keys (Jaccard matches)
    assert mi[0].tolist() == [9, 5, 7, int(NONE)] and mi[1].tolist() == [2, 8, 4, int(NONE)]
    fs = np.full((parts, nq, k), -np.inf, np.float32)
    fs[0, 0, :2], fs[2, 0, :1], fs[1, 1, :3] = [0.9, 0.1], [0.9], [0.5, 0.5, -0.2]
    mf = np.zeros((nq, k), np.float32)
    ctx.merge_topk_f32(ids, fs, parts, nq, k, mi, mf)
    assert mi[0].tolist() == [5, 7, 9, int(NONE)] and mi[1].tolist() == [2, 4, 8, int(NONE)]
    assert np.isneginf(mf[0, 3]) and mf[1, 2] == np.float32(-0.2)
