"""Parity of the CUDA Hamming top-k (through the C ABI) against the CPU oracle.

The reference has no Hamming scan; semantics are docs/HASH_SPEC.md section 6:
dist = popcount(q ^ c), total order (dist asc, record_id asc).  Integer work: bit-exact.
"""
import numpy as np
import pytest

import oracle
from ucfp_b200 import Context, Corpus, UcfpError, _ffi

pytestmark = pytest.mark.gpu
U64 = np.uint64


def _scan(ctx, codes, queries, k, ids=None, id_base=0):
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, max(len(codes), 1))
    if id_base:
        corpus.set_id_base(id_base)
    if len(codes):
        corpus.append(np.ascontiguousarray(codes, dtype=U64), None if ids is None else np.ascontiguousarray(ids, dtype=U64))
    got = corpus.scan_hamming(np.ascontiguousarray(queries, dtype=U64), k)
    corpus.close()
    return got


def _check(ctx, codes, queries, k, ids=None, id_base=0):
    gi, gd = _scan(ctx, codes, queries, k, ids, id_base)
    oi, od = oracle.hamming_topk(codes, queries, k, ids=ids, id_base=id_base, threads=oracle.host_threads())
    np.testing.assert_array_equal(gd, od)
    np.testing.assert_array_equal(gi, oi)


@pytest.mark.parametrize("n,nq,k", [(1, 1, 1), (7, 3, 10), (2047, 5, 10), (2048, 5, 10), (2049, 5, 10), (4097, 2, 1),
                                    (100_003, 64, 10), (262_144, 33, 100), (1_000_001, 17, 10),
                                    (5_000, 3, 1500), (200_000, 70, 1100), (1_030, 2, 2048)])
def test_random_corpus_matches_oracle(ctx, n, nq, k):
    codes = oracle.fill_u64(n, 0xC0DE)
    queries = oracle.fill_u64(nq, 0xBEEF)
    _check(ctx, codes, queries, k)
    assert ctx.last_scan_fallbacks() == 0      # the threshold path, not the exact fallback, produced this


def test_explicit_ids_and_id_base(ctx):
    n = 300_000
    codes = oracle.fill_u64(n, 1)
    rng = np.random.default_rng(7)
    ids = rng.permutation(n).astype(U64) * U64(1_000_003) + U64(5)
    queries = oracle.fill_u64(40, 2)
    _check(ctx, codes, queries, 10, ids=ids)
    _check(ctx, codes, queries, 10, id_base=10**12)


def test_planted_neighbours_and_heavy_ties(ctx):
    """integer distances tie constantly; the (dist, id) order must hold at the boundary."""
    n, nq, k = 500_000, 32, 10
    rng = np.random.default_rng(3)
    # only 6 distinct random bits per code -> thousands of exact ties for every query
    codes = rng.integers(0, 64, n).astype(U64) << U64(20)
    queries = rng.integers(0, 64, nq).astype(U64) << U64(20)
    for j in range(nq):  # planted neighbours at distance 0..11 (BASELINE config 2 construction)
        for d in range(12):
            mask = U64(0)
            for b in rng.choice(64, d, replace=False):
                mask |= U64(1) << U64(int(b))
            codes[rng.integers(0, n)] = queries[j] ^ mask
    ids = rng.permutation(n).astype(U64)
    _check(ctx, codes, queries, k, ids=ids)
    _check(ctx, codes, queries, 64, ids=ids)


def test_duplicate_flood_is_rescanned(ctx):
    """All rows identical and ids descending: every row beats the current k-th and the candidate lists overflow.  The flagged
    queries are scanned again under the bound their (sampled) truncated lists produced and must return the k smallest ids
    without reaching the exact multi-pass selection."""
    n, k = 300_000, 10
    codes = np.full(n, 0x0123456789ABCDEF, dtype=U64)
    ids = np.arange(n, 0, -1, dtype=U64) * U64(3)
    queries = np.array([0x0123456789ABCDEF, 0x0123456789ABCDEE, 0], dtype=U64)
    _check(ctx, codes, queries, k, ids=ids)
    assert (ctx.last_scan_fallbacks(), ctx.last_scan_exact_selects()) == (3, 0)
    codes[::2] ^= U64(1)  # two tie classes
    _check(ctx, codes, queries, 33, ids=ids)
    assert (ctx.last_scan_fallbacks(), ctx.last_scan_exact_selects()) == (3, 0)


def test_duplicate_flood_with_large_k(ctx):
    """k = 1000 of a 4096-entry list: every re-scan round shrinks the flood by only ~cap / 2k; whichever path ends up settling the
    queries (a second round or the exact selection), the k smallest ids must come back.  (The exact-selection kernel itself is
    pinned by tests/test_jaccard_gpu.py::test_duplicate_flood_without_rescan_rounds_takes_exact_fallback.)"""
    n, k = 300_000, 1000
    codes = np.full(n, 0x0123456789ABCDEF, dtype=U64)
    codes[::2] ^= U64(1)
    ids = np.arange(n, 0, -1, dtype=U64) * U64(3)
    queries = np.array([0x0123456789ABCDEF, 0], dtype=U64)
    _check(ctx, codes, queries, k, ids=ids)
    assert ctx.last_scan_fallbacks() == 2


# ---- batches of >= 16 queries run chunks of >= 2^19 rows on the int8 tensor pipe (hamming_mma_scan_kernel) ----------
@pytest.mark.parametrize("n,nq,k", [(700_001, 16, 10), (1_500_000, 130, 10), (2_000_003, 1024, 10), (1_250_000, 1000, 3),
                                    (900_000, 1025, 10), (1_100_000, 257, 100), (1_000_000, 640, 10), (800_000, 100, 10)])
def test_tensor_path_matches_oracle(ctx, n, nq, k):
    """Ragged query counts (partial 128-query tiles, > 1024 -> two passes) and ragged row counts (odd tails,
    partial 512-code tiles) through the tensor-core filter: bit-exact ids and distances."""
    codes = oracle.fill_u64(n, 0xC0DE + n)
    queries = oracle.fill_u64(nq, 0xBEEF + nq)
    queries[: min(nq, 8)] = codes[n - 1 - np.arange(min(nq, 8)) * 3]      # exact matches in the last (partial) tile
    _check(ctx, codes, queries, k)
    assert ctx.last_scan_fallbacks() == 0


def test_tensor_path_extreme_codes_and_distances(ctx):
    """All-zero / all-one codes and queries: x = +-64 is where the packed accumulator fields alias (dist 0 vs 64)."""
    n = 1_200_000
    codes = oracle.fill_u64(n, 5)
    codes[600_000::1000] = U64(0)
    codes[600_001::1000] = U64(2**64 - 1)
    codes[600_002::1000] = U64(0x00000000FFFFFFFF)
    queries = np.concatenate([np.array([0, 2**64 - 1, 0xFFFFFFFF00000000, 1, 2**63], dtype=U64), oracle.fill_u64(27, 6)])
    ids = np.arange(n, dtype=U64)[::-1].copy()
    _check(ctx, codes, queries, 10)
    _check(ctx, codes, queries, 50, ids=ids)


def test_tensor_path_heavy_ties_explicit_ids(ctx):
    """Thousands of exact ties per query inside the tensor-scanned chunks: the (dist, id) admission rule with
    permuted explicit ids, and an id_base, must hold at the boundary."""
    n, nq, k = 1_600_000, 64, 10
    rng = np.random.default_rng(11)
    codes = rng.integers(0, 4096, n).astype(U64) << U64(13)
    queries = rng.integers(0, 4096, nq).astype(U64) << U64(13)
    for j in range(nq):
        for d in range(12):
            mask = U64(0)
            for b in rng.choice(64, d, replace=False):
                mask |= U64(1) << U64(int(b))
            codes[rng.integers(n // 2, n)] = queries[j] ^ mask
    ids = rng.permutation(n).astype(U64) * U64(7) + U64(1)
    _check(ctx, codes, queries, k, ids=ids)
    _check(ctx, codes, queries, k, id_base=2**40)


def test_tensor_path_duplicate_flood_is_rescanned(ctx):
    n, k = 1_300_000, 10
    codes = oracle.fill_u64(n, 77)
    codes[700_000:] = U64(0xFEEDFACECAFEBEEF)                       # 600 K identical rows, ids descending
    ids = np.arange(n, 0, -1, dtype=U64)
    queries = np.concatenate([np.array([0xFEEDFACECAFEBEEF, 0xFEEDFACECAFEBEEE], dtype=U64), oracle.fill_u64(30, 78)])
    _check(ctx, codes, queries, k, ids=ids)
    assert ctx.last_scan_fallbacks() > 0 and ctx.last_scan_exact_selects() == 0


def test_tensor_path_incremental_appends_rebuild_pair_rows(ctx):
    """Operand rows pair codes (2r, 2r+1): appends that start at an odd row must rebuild the shared row; in-place edits
    followed by refresh() must be seen by the tensor scan."""
    import torch
    n1, n2, n3, n4 = 333_333, 1, 250_001, 415_665
    codes = oracle.fill_u64(n1 + n2 + n3 + n4, 31)
    queries = oracle.fill_u64(96, 32)
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, len(codes))
    pos = 0
    for m in (n1, n2, n3, n4):
        corpus.append(codes[pos:pos + m].copy())
        pos += m
    gi, gd = corpus.scan_hamming(queries, 10)
    oi, od = oracle.hamming_topk(codes, queries, 10, threads=oracle.host_threads())
    np.testing.assert_array_equal(gd, od)
    np.testing.assert_array_equal(gi, oi)
    # overwrite two rows in place with exact copies of queries, one at an odd and one at an even row
    class _A:
        __cuda_array_interface__ = {"shape": (len(codes),), "typestr": "<i8", "data": (corpus.device_rows_ptr(), False), "version": 2}
    view = torch.as_tensor(_A(), device="cuda")
    codes[777_777] = queries[5]; codes[900_000] = queries[6]
    view[777_777] = int(queries[5].view(np.int64)); view[900_000] = int(queries[6].view(np.int64))
    torch.cuda.synchronize()
    corpus.refresh()
    gi, gd = corpus.scan_hamming(queries, 10)
    oi, od = oracle.hamming_topk(codes, queries, 10, threads=oracle.host_threads())
    np.testing.assert_array_equal(gd, od)
    np.testing.assert_array_equal(gi, oi)
    assert gd[5, 0] == 0 and gi[5, 0] == 777_777 and gd[6, 0] == 0 and gi[6, 0] == 900_000
    corpus.close()


@pytest.mark.parametrize("switch", ["UCFP_HAMMING_NO_OPS=1", "UCFP_HAMMING_STAGGER=0"])
def test_tensor_path_variants_in_a_fresh_process(switch):
    """The scan also runs without the pre-built operand rows (their allocation is optional): the producer warps then expand
    the codes in the kernel (NO_OPS).  STAGGER=0 selects the lock-step epilogue of the stage-image kernel (the product path is the
    two-group epilogue).  The switches are read once per process, hence the subprocess; 200 and 1000 queries = two and eight
    query tiles (an even and an odd share of the accumulator items per epilogue group)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = ("import numpy as np, oracle\n"
            "from ucfp_b200 import Context, Corpus, _ffi\n"
            "ctx = Context(0)\n"
            "codes = oracle.fill_u64(1_234_567, 41); q = oracle.fill_u64(200, 42); q[:4] = codes[[5, 700_001, 1_234_566, 99_999]]\n"
            "c = Corpus(ctx, _ffi.KIND_HAMMING64, len(codes)); c.append(codes)\n"
            "gi, gd = c.scan_hamming(q, 10)\n"
            "oi, od = oracle.hamming_topk(codes, q, 10, threads=oracle.host_threads())\n"
            "assert (gi == oi).all() and (gd == od).all() and ctx.last_scan_fallbacks() == 0\n"
            "q = oracle.fill_u64(1000, 43); q[:3] = codes[[7, 600_000, 1_234_560]]\n"
            "gi, gd = c.scan_hamming(q, 10)\n"
            "oi, od = oracle.hamming_topk(codes, q, 10, threads=oracle.host_threads())\n"
            "assert (gi == oi).all() and (gd == od).all() and ctx.last_scan_fallbacks() == 0\n"
            "print('ok')\n")
    name, value = switch.split("=")
    env = dict(os.environ, PYTHONPATH=root)
    env[name] = value
    out = subprocess.run([sys.executable, "-c", prog], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_fewer_rows_than_k_pads_with_sentinels(ctx):
    codes = oracle.fill_u64(5, 9)
    queries = oracle.fill_u64(3, 10)
    gi, gd = _scan(ctx, codes, queries, 8)
    assert (gi[:, 5:] == U64(_ffi.ID_NONE)).all() and (gd[:, 5:] == np.uint32(2**32 - 1)).all()
    _check(ctx, codes, queries, 8)
    gi, gd = _scan(ctx, codes[:0], queries, 4)  # empty corpus
    assert (gi == U64(_ffi.ID_NONE)).all()


def test_nq_zero_and_k_zero_are_noops(ctx):
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, 16)
    corpus.append(oracle.fill_u64(16, 1))
    ids = np.full((0, 10), 7, dtype=U64)
    corpus.scan_hamming(np.zeros(0, dtype=U64), 10, ids, np.zeros((0, 10), dtype=np.uint32))
    corpus.close()


def test_wrong_kind_and_capacity_errors(ctx):
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, 8)
    with pytest.raises(UcfpError) as e:
        corpus.append(oracle.fill_u64(9, 1))
    assert e.value.code == _ffi.E_CAPACITY
    with pytest.raises(UcfpError) as e:
        corpus.scan_jaccard(np.zeros((1, 128), dtype=U64), 1)
    assert e.value.code == _ffi.E_STATE
    corpus.append(oracle.fill_u64(4, 1))
    with pytest.raises(UcfpError) as e:  # explicit ids after implicit ones
        corpus.append(oracle.fill_u64(2, 1), np.arange(2, dtype=U64))
    assert e.value.code == _ffi.E_STATE
    corpus.close()


def test_device_resident_buffers_and_synthetic_rows(ctx):
    """Inputs/outputs already in HBM (torch tensors): no staging, results identical to host buffers."""
    import torch
    n, nq, k = 3_000_000, 128, 10
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, n)
    corpus.append_synthetic(0xC0DE, 0, n)                 # device generator == oracle.fill_u64
    queries = oracle.fill_u64(nq, 0xBEEF)
    q_dev = torch.from_numpy(queries.view(np.int64)).cuda()
    ids_dev, dist_dev = corpus.scan_hamming(q_dev, k)
    torch.cuda.synchronize()
    oi, od = oracle.hamming_topk(oracle.fill_u64(n, 0xC0DE), queries, k, threads=oracle.host_threads())
    np.testing.assert_array_equal(ids_dev.cpu().numpy().view(U64), oi)
    np.testing.assert_array_equal(dist_dev.cpu().numpy().view(np.uint32), od)
    hi, hd = corpus.scan_hamming(queries, k)              # host buffers through the same corpus
    np.testing.assert_array_equal(hi, oi)
    np.testing.assert_array_equal(hd, od)
    corpus.close()


def test_config2_scale_100m_codes_1024_queries(ctx):
    """BASELINE config 2 at full size: 100 M synthetic codes, 1024 queries, k = 10, with planted
    neighbours.  Checked by size-independent properties plus the oracle on the planted structure."""
    import torch
    n, nq, k = 100_000_000, 1024, 10
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, n)
    corpus.append_synthetic(0xC0DE, 0, n)
    queries = oracle.fill_u64(nq, 0xBEEF)
    # plant, for every query j, neighbours at distance 0..11 at pseudo-random rows (written on device)
    rows = torch.from_numpy((oracle.fill_u64(nq * 12, 0xFACE) % U64(n)).astype(np.int64))
    planted = np.empty(nq * 12, dtype=U64)
    for j in range(nq):
        for d in range(12):
            planted[j * 12 + d] = queries[j] ^ U64((1 << d) - 1)
    dev_rows = corpus.device_rows_ptr()

    class _Arr:  # minimal __cuda_array_interface__ wrapper over the corpus rows
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (dev_rows, False), "version": 2}
    view = torch.as_tensor(_Arr(), device="cuda")
    uniq_rows, first = np.unique(rows.numpy(), return_index=True)   # colliding rows keep the first write
    view[torch.from_numpy(uniq_rows).cuda()] = torch.from_numpy(planted[first].view(np.int64)).cuda()
    torch.cuda.synchronize()
    corpus.refresh()                                      # rows were written in place: rebuild the operand rows
    q_dev = torch.from_numpy(queries.view(np.int64)).cuda()
    ids_dev, dist_dev = corpus.scan_hamming(q_dev, k)
    torch.cuda.synchronize()
    ids = ids_dev.cpu().numpy().view(U64)
    dist = dist_dev.cpu().numpy().view(np.uint32)
    # 1. ordered by (dist, id)
    assert (np.diff(dist.astype(np.int64), axis=1) >= 0).all()
    same = np.diff(dist.astype(np.int64), axis=1) == 0
    assert (np.diff(ids.astype(np.int64), axis=1)[same] > 0).all()
    # 2. every reported distance is the true distance of that row
    got_codes = view[torch.from_numpy(ids.astype(np.int64).ravel()).cuda()].cpu().numpy().view(U64).reshape(nq, k)
    true = np.array([[bin(int(c) ^ int(q)).count("1") for c in row] for row, q in zip(got_codes, queries)], dtype=np.uint32)
    np.testing.assert_array_equal(dist, true)
    # 3. the planted row at distance d (when it survived collisions) must appear whenever d < dist_k
    planted_at = {int(r): int(planted[i]) for r, i in zip(uniq_rows, first)}
    for j in range(0, nq, 7):
        for d in range(12):
            r = int(rows[j * 12 + d])
            if planted_at.get(r) == int(planted[j * 12 + d]) and d < dist[j, -1]:
                assert r in ids[j], (j, d)
    # 4. bit-exact agreement with the oracle over the full 100 M rows for a slice of the queries
    host_codes = view.cpu().numpy().view(U64)
    sel = np.arange(0, nq, 16)
    oi, od = oracle.hamming_topk(host_codes, queries[sel], k, threads=oracle.host_threads())
    np.testing.assert_array_equal(ids[sel], oi)
    np.testing.assert_array_equal(dist[sel], od)
    corpus.close()
