"""IndexBackend::upsert / ::delete on the HBM mirror (src/index/mod.rs:20-25: "Insert-or-replace by (tenant_id,
record_id)", delete "Idempotent -- missing IDs are not an error").  After every mutation the corpus must answer
exactly like the oracle run over the same live (id, row) set -- for all three scans, on the POPC and on the
tensor-core paths -- because the side arrays of moved / replaced rows are re-derived in place."""
import numpy as np
import pytest

import oracle
from ucfp_b200 import Corpus, UcfpError, _ffi

pytestmark = pytest.mark.gpu
U64 = np.uint64


class Model:
    """host-side truth: dict id -> row, answered by the oracle"""

    def __init__(self):
        self.rows = {}

    def upsert(self, ids, rows):
        for i, r in zip(ids, rows):
            self.rows[int(i)] = np.array(r, copy=True)

    def delete(self, ids):
        return sum(self.rows.pop(int(i), None) is not None for i in set(int(x) for x in ids))

    def arrays(self):
        ids = np.array(sorted(self.rows), dtype=U64)
        rows = np.stack([self.rows[int(i)] for i in ids]) if len(ids) else np.zeros((0,), dtype=U64)
        return ids, rows


def _check_hamming(corpus, model, queries, k):
    ids, rows = model.arrays()
    gi, gd = corpus.scan_hamming(queries, k)
    oi, od = oracle.hamming_topk(rows.reshape(-1), queries, k, ids=ids, threads=oracle.host_threads())
    np.testing.assert_array_equal(gd, od)
    np.testing.assert_array_equal(gi, oi)


@pytest.mark.parametrize("n,nq", [(3_000, 5), (300_000, 96)])   # POPC scan only / tensor-core scan for the large chunks
def test_hamming_upsert_delete_replace(ctx, n, nq):
    rng = np.random.default_rng(n)
    model = Model()
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, 1024)            # far too small: upsert grows it
    ids = rng.permutation(np.arange(10 * n, dtype=U64))[:n] * U64(7919)
    codes = oracle.fill_u64(n, 11)
    queries = oracle.fill_u64(nq, 12)
    codes[: nq] = queries ^ U64(1)                              # a near neighbour per query among the first rows
    assert corpus.upsert(ids, codes) == 0
    model.upsert(ids, codes.reshape(-1, 1))
    assert len(corpus) == n and corpus.allocated >= n
    _check_hamming(corpus, model, queries, 10)
    # replace a third of the rows in place (some twice in one batch: the last occurrence wins), insert new ones
    rep = ids[rng.choice(n, n // 3, replace=False)]
    batch_ids = np.concatenate([rep, rep[:50], np.arange(5, dtype=U64) + U64(3)])
    batch_rows = oracle.fill_u64(len(batch_ids), 13)
    batch_rows[:nq] = queries                                   # exact duplicates of the queries replace old rows
    replaced = corpus.upsert(batch_ids, batch_rows)
    model.upsert(batch_ids, batch_rows.reshape(-1, 1))
    assert replaced == len(rep) and len(corpus) == len(model.rows)
    _check_hamming(corpus, model, queries, 10)
    # delete: half of the rows incl. the near neighbours, unknown ids, duplicates in the list
    victims = np.concatenate([ids[::2], ids[:10], np.array([1, 2], dtype=U64)])
    removed = corpus.delete(victims)
    assert removed == model.delete(victims) and len(corpus) == len(model.rows)
    _check_hamming(corpus, model, queries, 10)
    assert corpus.delete(victims) == 0                          # idempotent
    # delete everything, then the corpus is usable again
    all_ids, _ = model.arrays()
    assert corpus.delete(all_ids) == len(all_ids) and len(corpus) == 0
    gi, gd = corpus.scan_hamming(queries, 3)
    assert (gi == _ffi.ID_NONE).all() and (gd == 2**32 - 1).all()
    corpus.upsert(np.array([42], dtype=U64), queries[:1].copy())
    gi, gd = corpus.scan_hamming(queries[:1], 2)
    assert gi[0, 0] == 42 and gd[0, 0] == 0 and gi[0, 1] == _ffi.ID_NONE
    corpus.close()


def test_delete_on_an_implicit_id_corpus_materialises_ids(ctx):
    n, nq = 200_000, 80
    codes = oracle.fill_u64(n, 21)
    queries = oracle.fill_u64(nq, 22)
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, n)
    corpus.set_id_base(1_000_000)
    corpus.append(codes)
    model = Model()
    model.upsert(np.arange(n, dtype=U64) + U64(1_000_000), codes.reshape(-1, 1))
    gi, _ = corpus.scan_hamming(queries, 10)
    victims = np.unique(gi[:, :3].reshape(-1))                  # every query loses its three best hits
    assert corpus.delete(victims) == model.delete(victims) == len(victims)
    _check_hamming(corpus, model, queries, 10)
    corpus.close()


def test_jaccard_upsert_delete(ctx):
    rng = np.random.default_rng(5)
    n, nq = 30_000, 6
    sigs = oracle.fill_u64(n * 128, 31).reshape(n, 128)
    q = oracle.fill_u64(nq * 128, 32).reshape(nq, 128)
    for j in range(nq):
        sigs[100 + j, : 60 + 10 * j] = q[j, : 60 + 10 * j]
    ids = (np.arange(n, dtype=U64) * U64(3) + U64(17))
    corpus = Corpus(ctx, _ffi.KIND_MINHASH128, 256)
    model = Model()
    corpus.upsert(ids, sigs); model.upsert(ids, sigs)

    def check():
        mi, mr = model.arrays()
        gi, gm = corpus.scan_jaccard(q, 10)
        oi, om = oracle.jaccard_topk(mr, q, 10, ids=mi, threads=oracle.host_threads())
        np.testing.assert_array_equal(gm, om)
        np.testing.assert_array_equal(gi, oi)

    check()
    victims = np.concatenate([ids[100:103], ids[rng.choice(n, 5_000, replace=False)]])
    assert corpus.delete(victims) == model.delete(victims)
    check()
    new_rows = oracle.fill_u64(40 * 128, 33).reshape(40, 128)
    new_rows[0, :120] = q[0, :120]
    new_ids = np.concatenate([ids[200:220], np.arange(20, dtype=U64) + U64(10**9)])
    corpus.upsert(new_ids, new_rows); model.upsert(new_ids, new_rows)
    check()
    corpus.close()


def test_cosine_upsert_delete(ctx):
    rng = np.random.default_rng(9)
    n, dim, nq = 40_000, 96, 5
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    rows[7] = q[0] * 3.0
    ids = np.arange(n, dtype=U64) + U64(500)
    corpus = Corpus(ctx, _ffi.KIND_COSINE, 128, dim=dim)
    model = Model()
    corpus.upsert(ids, rows); model.upsert(ids, rows)

    def check():
        mi, mr = model.arrays()
        gi, gs = corpus.scan_cosine(q, 10)
        oi, os_, _ = oracle.cosine_topk(mr, q, 10, ids=mi, mode=1, threads=oracle.host_threads())
        np.testing.assert_array_equal(gi, oi)
        np.testing.assert_array_equal(gs.view(np.uint32), os_.view(np.uint32))

    check()
    assert corpus.delete(ids[:2000]) == model.delete(ids[:2000]) == 2000      # drops the planted best hit of query 0
    check()
    rep = rng.standard_normal((300, dim)).astype(np.float32)
    rep[5] = q[1] * 0.5
    rid = np.concatenate([ids[3000:3200], np.arange(100, dtype=U64)])
    corpus.upsert(rid, rep); model.upsert(rid, rep)
    check()
    corpus.close()


def test_reserve_keeps_rows_and_rejects_nothing(ctx):
    codes = oracle.fill_u64(70_000, 41)
    q = oracle.fill_u64(70, 42)
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, 70_000)
    corpus.append(codes)
    before = corpus.scan_hamming(q, 10)
    corpus.reserve(1_000_000)
    assert corpus.allocated >= 1_000_000 and len(corpus) == 70_000
    after = corpus.scan_hamming(q, 10)
    np.testing.assert_array_equal(before[0], after[0])
    np.testing.assert_array_equal(before[1], after[1])
    with pytest.raises(UcfpError):
        corpus.upsert(np.array([_ffi.ID_NONE], dtype=U64), q[:1].copy())   # the sentinel is not a record id
    corpus.close()
