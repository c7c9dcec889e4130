"""Parity of the CUDA cosine top-k (tcgen05 coarse pass + exact f32 rescoring, through the C ABI) against the
CPU restatement of EmbeddedBackend::knn (reference src/index/embedded/mod.rs:268-360, :454-495).

The GPU path returns the reference's own f32 scores (same summation order, no FMA), so scores are compared
bit-for-bit -- far inside the 1e-4 relative tolerance BASELINE.json allows -- and ids must be identical under
the total order (score desc, record_id asc).  The reference's tie behaviour is insertion-order dependent
(SURVEY A7); with distinct scores both orders coincide, which the mode-0 comparisons check."""
import numpy as np
import pytest

import oracle
from ucfp_b200 import Corpus, _ffi

pytestmark = pytest.mark.gpu
U64 = np.uint64


def _scan(ctx, rows, q, k, ids=None, id_base=0):
    n, dim = rows.shape
    corpus = Corpus(ctx, _ffi.KIND_COSINE, max(n, 1), dim=dim)
    if id_base:
        corpus.set_id_base(id_base)
    if n:
        corpus.append(np.ascontiguousarray(rows, dtype=np.float32), None if ids is None else np.ascontiguousarray(ids, dtype=U64))
    gi, gs = corpus.scan_cosine(np.ascontiguousarray(q, dtype=np.float32), k)
    corpus.close()
    return gi, gs


def _check(ctx, rows, q, k, ids=None, id_base=0, also_reference_order=True):
    gi, gs = _scan(ctx, rows, q, k, ids, id_base)
    oi, osc, cnt = oracle.cosine_topk(rows, q, k, ids=ids, id_base=id_base, mode=1, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gs.view(np.uint32), osc.view(np.uint32))      # bit-exact scores
    valid = oi != U64(_ffi.ID_NONE)
    if valid.any():
        rel = np.abs(gs[valid] - osc[valid]) / np.maximum(np.abs(osc[valid]), 1e-30)
        assert rel.max() <= 1e-4                                                  # the stated tolerance
    if also_reference_order:
        ri, rs, _ = oracle.cosine_topk(rows, q, k, ids=ids, id_base=id_base, mode=0, threads=1)
        np.testing.assert_array_equal(gi, ri)


def unit_rows(n, dim, seed, bf16_exact=False):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if bf16_exact:  # BASELINE config 4: values representable in bf16
        x = (x.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    return x


def test_reference_known_answer_round_trip(ctx):
    """embedded/mod.rs:523-544 upsert_and_knn_round_trip."""
    rows = np.array([[1, 0, 0], [0, 1, 0], [0.7, 0.7, 0]], np.float32)
    gi, gs = _scan(ctx, rows, np.array([[0.6, 0.6, 0.0]], np.float32), 2, ids=[100, 200, 300])
    assert gi[0, 0] == 300 and gs[0, 0] > gs[0, 1]
    _check(ctx, rows, np.array([[0.6, 0.6, 0.0]], np.float32), 2, ids=[100, 200, 300], also_reference_order=False)


def test_reference_known_answer_server_query(ctx):
    """server/tests.rs:53-113 upsert_then_query_round_trips: the record closest to the query wins."""
    rows = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0]], np.float32)
    gi, gs = _scan(ctx, rows, np.array([[0.1, 0.9, 0.0, 0.0]], np.float32), 10, ids=[100, 200, 300])
    assert gi[0, 0] == 200 and (gi[0, 3:] == U64(_ffi.ID_NONE)).all() and np.isneginf(gs[0, 3:]).all()


def test_zero_norm_rows_and_queries_never_match(ctx):
    """embedded/mod.rs:284 (zero query -> empty) and :328-330 (zero rows skipped)."""
    rows = unit_rows(3000, 64, 1)
    rows[::7] = 0.0
    q = unit_rows(4, 64, 2)
    q[2] = 0.0
    gi, gs = _scan(ctx, rows, q, 10)
    assert (gi[2] == U64(_ffi.ID_NONE)).all()
    assert not np.isin(gi[[0, 1, 3]], np.arange(0, 3000, 7).astype(U64)).any()
    _check(ctx, rows, q, 10)


@pytest.mark.parametrize("n,dim,nq,k", [(1, 8, 1, 1), (5, 3, 2, 10), (1023, 512, 3, 10), (1024, 512, 3, 10), (1025, 512, 3, 10),
                                        (5000, 100, 17, 10), (20_000, 768, 33, 10), (100_000, 512, 64, 10), (60_000, 512, 300, 25),
                                        (9_000, 1000, 5, 100)])
def test_random_corpus_matches_oracle(ctx, n, dim, nq, k):
    rows = unit_rows(n, dim, n + dim) * np.float32(3.7)      # norms != 1 on purpose
    q = unit_rows(nq, dim, 5) * np.float32(0.2)
    _check(ctx, rows, q, k)
    assert ctx.last_scan_fallbacks() == 0      # the threshold path, not the exact fallback, produced this


def test_planted_neighbours_bf16_exact_corpus(ctx):
    """BASELINE config 4 construction: bf16-representable unit rows, 8 planted neighbours per query."""
    n, dim, nq, k = 200_000, 512, 128, 10
    rows = unit_rows(n, dim, 11, bf16_exact=True)
    q = unit_rows(nq, dim, 12, bf16_exact=True)
    rng = np.random.default_rng(13)
    for j in range(nq):
        for t in range(8):
            v = q[j] + np.float32(0.05 * (t + 1)) * rng.standard_normal(dim).astype(np.float32)
            rows[rng.integers(0, n)] = v / np.linalg.norm(v)
    ids = rng.permutation(n).astype(U64) + U64(10**10)
    _check(ctx, rows, q, k, ids=ids)


def test_exact_ties_are_ordered_by_id(ctx):
    """Duplicated rows give bit-identical scores: the GPU order is (score desc, id asc); the reference's
    membership among boundary ties depends on scan order (documented), so only mode 1 is compared."""
    rows = np.repeat(unit_rows(50, 64, 3), 40, axis=0)        # 2000 rows, 40 copies of each vector
    ids = np.random.default_rng(4).permutation(len(rows)).astype(U64)
    _check(ctx, rows, unit_rows(6, 64, 5), 10, ids=ids, also_reference_order=False)


def test_candidate_overflow_takes_exact_fallback(ctx):
    """More than kCap rows inside the coarse margin of the k-th score: flagged queries are recomputed by the
    cooperative exact selection."""
    n, dim = 30_000, 64
    base = unit_rows(1, dim, 7)[0]
    rng = np.random.default_rng(8)
    rows = (base[None, :] + 1e-4 * rng.standard_normal((n, dim))).astype(np.float32)   # all within 1e-4 of each other
    ids = np.arange(n, 0, -1, dtype=U64)
    _check(ctx, rows, np.stack([base, -base]), 10, ids=ids, also_reference_order=False)
    assert ctx.last_scan_fallbacks() > 0       # this input must have gone through the exact selection


def test_device_buffers(ctx):
    import torch
    n, dim, nq, k = 50_000, 512, 40, 10
    rows, q = unit_rows(n, dim, 21), unit_rows(nq, dim, 22)
    corpus = Corpus(ctx, _ffi.KIND_COSINE, n, dim=dim)
    corpus.append(torch.from_numpy(rows).cuda())
    gi, gs = corpus.scan_cosine(torch.from_numpy(q).cuda(), k)
    torch.cuda.synchronize()
    oi, osc, _ = oracle.cosine_topk(rows, q, k, mode=1, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi.cpu().numpy().view(U64), oi)
    np.testing.assert_array_equal(gs.cpu().numpy().view(np.uint32), osc.view(np.uint32))
    corpus.close()
