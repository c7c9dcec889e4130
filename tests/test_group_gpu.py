"""Multi-GPU through the boundary (SURVEY 8e process model: a single process driving the GPUs of one box with
ncclCommInitAll).  ucfp_group_scan_* over record-range shards must be byte-identical to the single-corpus scan and
to the oracle, with and without the exchange of admission bounds between the ranks while they walk their shards.  Uses as many
GPUs as the box has (1, 2, 4 or 8); with one GPU the group degenerates to a plain scan and only the plumbing is checked."""
import numpy as np
import pytest

import oracle
from ucfp_b200 import Corpus, Group, _ffi
from ucfp_b200.sharding import shard_range

pytestmark = pytest.mark.gpu
U64 = np.uint64


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module", params=[0, 4], ids=["no-bound-exchange", "4-bound-exchanges"])
def group(request):
    """The default group filters every shard at its own bound; UCFP_GROUP_EXCHANGES (read at group creation) turns on the
    cross-rank bound exchange.  Both must give the same bytes."""
    import os
    os.environ["UCFP_GROUP_EXCHANGES"] = str(request.param)
    g = Group.local(list(range(min(_n_gpus(), 8))))
    del os.environ["UCFP_GROUP_EXCHANGES"]
    yield g
    g.close()


def _shards(group, kind, rows, dim=0, explicit_ids=None):
    out, world = [], group.local_size
    for r in range(world):
        lo, hi = shard_range(len(rows), r, world)
        c = Corpus(group.ctx(r), kind, max(hi - lo, 1), dim=dim)
        if explicit_ids is None:
            c.set_id_base(lo)
            if hi > lo:
                c.append(np.ascontiguousarray(rows[lo:hi]))
        elif hi > lo:
            c.append(np.ascontiguousarray(rows[lo:hi]), np.ascontiguousarray(explicit_ids[lo:hi]))
        out.append(c)
    return out


@pytest.mark.parametrize("n,nq,k", [(50, 3, 10), (3_000_000, 200, 10), (40_000_000, 1024, 10), (700_000, 5, 100)])
def test_group_hamming_equals_single_corpus_and_oracle(group, n, nq, k):
    codes = oracle.fill_u64(n, 0xC0DE)
    queries = oracle.fill_u64(nq, 0xBEEF)
    rng = np.random.default_rng(n)
    for j in range(min(nq, 64)):                       # planted neighbours at distances 0..11 anywhere in the corpus
        for d in range(12):
            codes[rng.integers(0, n)] = queries[j] ^ U64((1 << d) - 1)
    shards = _shards(group, _ffi.KIND_HAMMING64, codes)
    gi, gd = group.scan_hamming(shards, queries, k)
    oi, od = oracle.hamming_topk(codes, queries, k, threads=oracle.host_threads())
    np.testing.assert_array_equal(gd, od)
    np.testing.assert_array_equal(gi, oi)
    single = Corpus(group.ctx(0), _ffi.KIND_HAMMING64, n)
    single.append(codes)
    si, sd = single.scan_hamming(queries, k)
    np.testing.assert_array_equal(gi, si)
    np.testing.assert_array_equal(gd, sd)
    for c in shards + [single]:
        c.close()


def test_group_hamming_explicit_ids_heavy_ties_and_device_buffers(group):
    import torch
    n, nq, k = 2_000_000, 96, 10
    rng = np.random.default_rng(1)
    codes = oracle.fill_u64(n, 5)
    queries = oracle.fill_u64(nq, 6)
    codes[rng.choice(n, n // 50, replace=False)] = queries[0] ^ U64(3)      # 40 000 rows tie at distance 2 for query 0
    ids = rng.permutation(n).astype(U64) * U64(13) + U64(5)
    shards = _shards(group, _ffi.KIND_HAMMING64, codes, explicit_ids=ids)
    dev = torch.device("cuda", group.devices[-1])                              # queries and outputs on the LAST local GPU
    qd = torch.from_numpy(queries.view(np.int64)).to(dev)
    gi, gd = group.scan_hamming(shards, qd, k)
    torch.cuda.synchronize(dev)
    oi, od = oracle.hamming_topk(codes, queries, k, ids=ids, threads=oracle.host_threads())
    np.testing.assert_array_equal(gd.cpu().numpy().view(np.uint32), od)
    np.testing.assert_array_equal(gi.cpu().numpy().view(np.uint64), oi)
    for c in shards:
        c.close()


def test_group_jaccard_and_cosine(group):
    rng = np.random.default_rng(2)
    sig = oracle.fill_u64(60_000 * 128, 7).reshape(-1, 128)
    q = oracle.fill_u64(9 * 128, 8).reshape(-1, 128)
    sig[59_990, :110] = q[2, :110]
    sig[10, :64] = q[2, :64]
    shards = _shards(group, _ffi.KIND_MINHASH128, sig)
    gi, gm = group.scan_jaccard(shards, q, 10)
    oi, om = oracle.jaccard_topk(sig, q, 10, threads=oracle.host_threads())
    np.testing.assert_array_equal(gm, om)
    np.testing.assert_array_equal(gi, oi)
    [c.close() for c in shards]
    vec = rng.standard_normal((80_000, 128)).astype(np.float32)
    qv = rng.standard_normal((70, 128)).astype(np.float32)
    shards = _shards(group, _ffi.KIND_COSINE, vec, dim=128)
    gi, gs = group.scan_cosine(shards, qv, 10)
    oi, osc, _ = oracle.cosine_topk(vec, qv, 10, mode=1, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gs.view(np.uint32), osc.view(np.uint32))
    [c.close() for c in shards]
