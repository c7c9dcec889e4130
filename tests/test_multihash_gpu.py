"""block-hash-aware compare and re-rank over `multi` bundles (docs/HASH_SPEC.md section 10; the reference's
compare-time MultiHashConfig: src/modality/image.rs:21-24,90-104, src/server/dto.rs:462-480, defaults
web/src/lib/docs/api-reference-image.md:51-62).  f32 scores are bit-exact against the oracle: every operation is
separately rounded in a fixed order."""
import numpy as np
import pytest

import oracle
from ucfp_b200 import Corpus, UcfpError, _ffi

pytestmark = pytest.mark.gpu
U64 = np.uint64


def _bundles(n, seed):
    return oracle.fill_u64(n * 51, seed).reshape(n, 51)


def _near(b, rng, flips):
    """a bundle a few bit flips per word away from b (a re-encoded / slightly edited image)"""
    out = b.copy()
    for w in range(51):
        for bit in rng.choice(64, size=rng.integers(0, flips + 1), replace=False):
            out[w] ^= U64(1) << U64(bit)
    return out


def test_pairwise_compare_matches_the_oracle_bit_for_bit(ctx):
    rng = np.random.default_rng(0)
    a = _bundles(500, 1)
    b = np.stack([_near(x, rng, f) for x, f in zip(a, rng.integers(0, 20, size=500))])
    b[0] = a[0]
    for cfg in (None, {"block_distance_threshold": 0}, {"ahash_weight": 0.0, "phash_weight": 1.0, "dhash_weight": 0.0, "global_weight": 0.0},
                {"ahash_weight": 0.25, "phash_weight": 0.3, "dhash_weight": 0.45, "global_weight": 0.7, "block_weight": 0.2, "block_distance_threshold": 5},
                {"ahash_weight": 0.0, "phash_weight": 0.0, "dhash_weight": 0.0}, {"global_weight": 0.0, "block_weight": 0.0}):
        got = ctx.multihash_compare(a, b, cfg)
        want = np.array([oracle.multihash_score(x, y, cfg) for x, y in zip(a, b)], dtype=np.float32)
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    assert ctx.multihash_compare(a, b)[0] == 1.0
    with pytest.raises(UcfpError):
        ctx.multihash_compare(a, b, {"phash_weight": 1.5})


@pytest.mark.parametrize("n,nq,kp,k", [(5_000, 7, 50, 10), (300_000, 130, 64, 10), (40, 3, 100, 100)])
def test_rerank_matches_the_oracle(ctx, n, nq, kp, k):
    rng = np.random.default_rng(n)
    rows = _bundles(n, 3)
    queries = _bundles(nq, 4)
    if n >= 4 * nq:                                         # near-duplicates of every query at several edit strengths, at distinct rows
        spots = rng.choice(n, size=4 * nq, replace=False).reshape(nq, 4)
        for j in range(nq):
            for f, r in zip((0, 2, 6, 12), spots[j]):
                rows[r] = _near(queries[j], rng, f)
    ids = rng.permutation(10 * n)[:n].astype(U64) + U64(7)
    corpus = Corpus(ctx, _ffi.KIND_MULTIHASH, n)
    corpus.append(rows, ids)
    gi, gs = corpus.scan_multihash(queries, kp, k)
    oi, osc = oracle.multihash_rerank(rows, queries, kp, k, ids=ids, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gs.view(np.uint32), osc.view(np.uint32))
    if n > 100:
        assert (gs[:, 0] == 1.0).all()                      # the exact duplicate wins
    cfg = {"phash_weight": 0.1, "dhash_weight": 0.6, "block_weight": 0.9, "block_distance_threshold": 3}
    gi, gs = corpus.scan_multihash(queries, kp, min(k, 5), cfg)
    oi, osc = oracle.multihash_rerank(rows, queries, kp, min(k, 5), ids=ids, cfg=cfg, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gs.view(np.uint32), osc.view(np.uint32))
    corpus.close()


def test_hydration_from_536_byte_bundles_and_mutation(ctx):
    """Rows arrive as stored MultiHashFingerprint blobs (exact[32] | 3 x ImageFingerprint[168]); delete / upsert keep the
    PHash side corpus in step with the rows."""
    rng = np.random.default_rng(5)
    n, nq = 20_000, 70
    rows = _bundles(n, 6)
    queries = _bundles(nq, 7)
    for j in range(nq):
        rows[100 + j] = _near(queries[j], rng, 1)
    blobs = np.zeros((n, 536), dtype=np.uint8)
    for a in range(3):
        blobs[:, 32 + 168 * a + 32: 32 + 168 * a + 168] = rows[:, 17 * a: 17 * a + 17].copy().view(np.uint8).reshape(n, 136)
    ids = np.arange(n, dtype=U64) * U64(3)
    corpus = Corpus(ctx, _ffi.KIND_MULTIHASH, n)
    corpus.append_strided(blobs, 536, 0, n, ids)
    gi, gs = corpus.scan_multihash(queries, 32, 5)
    oi, osc = oracle.multihash_rerank(rows, queries, 32, 5, ids=ids, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gs.view(np.uint32), osc.view(np.uint32))
    # delete the planted best hits, replace a few rows, add new ones: answers follow
    victims = ids[100: 100 + nq: 2]
    assert corpus.delete(victims) == len(victims)
    keep = np.ones(n, dtype=bool); keep[100: 100 + nq: 2] = False
    rows2, ids2 = rows[keep].copy(), ids[keep].copy()
    new_rows = np.stack([_near(queries[j], rng, 0) for j in range(4)])
    new_ids = np.array([ids2[5], ids2[6], 10**9, 10**9 + 1], dtype=U64)
    corpus.upsert(new_ids, new_rows)
    rows2[5], rows2[6] = new_rows[0], new_rows[1]
    rows2 = np.concatenate([rows2, new_rows[2:]]); ids2 = np.concatenate([ids2, new_ids[2:]])
    gi, gs = corpus.scan_multihash(queries, 32, 5)
    oi, osc = oracle.multihash_rerank(rows2, queries, 32, 5, ids=ids2, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gs.view(np.uint32), osc.view(np.uint32))
    corpus.close()
