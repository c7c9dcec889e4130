"""Parity of the CUDA MinHash-128 Jaccard top-k (through the C ABI) against the CPU oracle: bit-exact ids and
match counts under the total order (matches desc, record_id asc).  Absent from the reference; the
signature layout is txtfp's MinHashSig<128> payload (src/modality/text.rs:200-204, server/tests.rs:1153-1162)."""
import numpy as np
import pytest

import oracle
from ucfp_b200 import Corpus, UcfpError, _ffi

pytestmark = pytest.mark.gpu
U64 = np.uint64


def synth(n, nq, seed, plant_frac=0.01):
    """BASELINE config 3: uniform slots; for plant_frac of the rows copy a random query's slots with
    probability p in {0.9, 0.7, 0.5} per slot."""
    sigs = oracle.fill_u64(n * 128, seed).reshape(n, 128)
    q = oracle.fill_u64(nq * 128, seed + 1).reshape(nq, 128)
    rng = np.random.default_rng(seed)
    rows = rng.choice(n, max(1, int(n * plant_frac)), replace=False)
    for r in rows:
        j = rng.integers(0, nq)
        p = rng.choice([0.9, 0.7, 0.5])
        mask = rng.random(128) < p
        sigs[r, mask] = q[j, mask]
    return sigs, q


def _check(ctx, sigs, q, k, ids=None, id_base=0):
    corpus = Corpus(ctx, _ffi.KIND_MINHASH128, max(len(sigs), 1))
    if id_base:
        corpus.set_id_base(id_base)
    if len(sigs):
        corpus.append(np.ascontiguousarray(sigs), None if ids is None else np.ascontiguousarray(ids, dtype=U64))
    gi, gm = corpus.scan_jaccard(np.ascontiguousarray(q), k)
    corpus.close()
    oi, om = oracle.jaccard_topk(sigs, q, k, ids=ids, id_base=id_base, threads=oracle.host_threads())
    np.testing.assert_array_equal(gm, om)
    np.testing.assert_array_equal(gi, oi)


@pytest.mark.parametrize("n,nq,k", [(1, 1, 1), (33, 3, 10), (1023, 4, 10), (1024, 4, 10), (1025, 4, 10), (9217, 5, 3),
                                    (50_000, 16, 10), (200_000, 64, 10), (300_001, 7, 100)])
def test_planted_corpus_matches_oracle(ctx, n, nq, k):
    sigs, q = synth(n, nq, 11 + n % 7)
    _check(ctx, sigs, q, k)
    assert ctx.last_scan_fallbacks() == 0      # the threshold path, not the exact fallback, produced this


def test_queries_without_any_neighbour(ctx):
    """No planted rows: all true matches are 0, so the answer is the k smallest ids -- the degenerate case in
    which byte collisions force verification of many rows."""
    n = 120_000
    sigs = oracle.fill_u64(n * 128, 5).reshape(n, 128)
    q = oracle.fill_u64(3 * 128, 6).reshape(3, 128)
    rng = np.random.default_rng(1)
    ids = rng.permutation(n).astype(U64) + U64(1000)
    _check(ctx, sigs, q, 10, ids=ids)
    _check(ctx, sigs, q, 10, id_base=7_000_000_000)


def test_byte_collisions_are_not_matches(ctx):
    """Rows that agree with the query in the low byte of every slot but differ above it: the sketch says 128,
    the truth says 0 -- they must lose to a genuine 3-slot match."""
    n = 5000
    sigs = oracle.fill_u64(n * 128, 9).reshape(n, 128)
    q = oracle.fill_u64(128, 10).reshape(1, 128)
    sigs[100:200] = (q[0] & U64(0xFF)) | (sigs[100:200] & ~U64(0xFF)) ^ U64(0x100)
    sigs[4000, :3] = q[0, :3]
    _check(ctx, sigs, q, 5)


def test_duplicate_flood_is_rescanned(ctx):
    n, k = 40_000, 10
    q = oracle.fill_u64(128, 3).reshape(1, 128)
    sigs = np.repeat(q, n, axis=0)
    sigs[::3, 5] ^= U64(1)  # two tie classes: 128 and 127 matches
    ids = np.arange(n, 0, -1, dtype=U64) * U64(5)
    _check(ctx, sigs, np.concatenate([q, q ^ U64(1)]), k, ids=ids)
    assert ctx.last_scan_fallbacks() > 0 and ctx.last_scan_exact_selects() == 0   # lists overflowed; the re-scan settled them


def test_duplicate_flood_without_rescan_rounds_takes_exact_fallback():
    """UCFP_RESCAN_ROUNDS=0 (read once per process, hence the subprocess) sends overflowed queries straight to the cooperative
    exact-selection kernel, which must return the same answer: the backstop for floods the re-scan rounds cannot shrink."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = ("import numpy as np, oracle\n"
            "from ucfp_b200 import Context, Corpus, _ffi\n"
            "U64 = np.uint64\n"
            "ctx = Context(0)\n"
            "n, k = 40_000, 10\n"
            "q = oracle.fill_u64(128, 3).reshape(1, 128)\n"
            "sigs = np.repeat(q, n, axis=0); sigs[::3, 5] ^= U64(1)\n"
            "ids = np.arange(n, 0, -1, dtype=U64) * U64(5)\n"
            "qq = np.concatenate([q, q ^ U64(1)])\n"
            "c = Corpus(ctx, _ffi.KIND_MINHASH128, n); c.append(sigs, ids)\n"
            "gi, gm = c.scan_jaccard(qq, k)\n"
            "oi, om = oracle.jaccard_topk(sigs, qq, k, ids=ids, threads=oracle.host_threads())\n"
            "assert (gi == oi).all() and (gm == om).all() and ctx.last_scan_exact_selects() > 0\n"
            "codes = np.full(300_000, 0x0123456789ABCDEF, dtype=U64); hid = np.arange(300_000, 0, -1, dtype=U64)\n"
            "h = Corpus(ctx, _ffi.KIND_HAMMING64, len(codes)); h.append(codes, hid)\n"
            "hq = np.array([0x0123456789ABCDEF, 7], dtype=U64)\n"
            "gi, gd = h.scan_hamming(hq, k)\n"
            "oi, od = oracle.hamming_topk(codes, hq, k, ids=hid, threads=oracle.host_threads())\n"
            "assert (gi == oi).all() and (gd == od).all() and ctx.last_scan_exact_selects() == 2\n"
            "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=root, UCFP_RESCAN_ROUNDS="0")
    out = subprocess.run([sys.executable, "-c", prog], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_fewer_rows_than_k_and_empty(ctx):
    sigs, q = synth(6, 2, 3, plant_frac=0.5)
    _check(ctx, sigs, q, 9)
    corpus = Corpus(ctx, _ffi.KIND_MINHASH128, 4)
    gi, gm = corpus.scan_jaccard(q, 3)
    assert (gi == U64(_ffi.ID_NONE)).all() and (gm == np.uint32(2**32 - 1)).all()
    with pytest.raises(UcfpError) as e:
        corpus.scan_hamming(np.zeros(1, dtype=U64), 1)
    assert e.value.code == _ffi.E_STATE
    corpus.close()


def test_device_buffers_and_incremental_append(ctx):
    import torch
    n, nq, k = 150_000, 32, 10
    sigs, q = synth(n, nq, 21)
    corpus = Corpus(ctx, _ffi.KIND_MINHASH128, n)
    for lo in range(0, n, 40_001):  # appends that do not align with sketch tiles
        corpus.append(torch.from_numpy(sigs[lo:lo + 40_001].view(np.int64)).cuda())
    gi, gm = corpus.scan_jaccard(torch.from_numpy(q.view(np.int64)).cuda(), k)
    torch.cuda.synchronize()
    oi, om = oracle.jaccard_topk(sigs, q, k, threads=oracle.host_threads())
    np.testing.assert_array_equal(gi.cpu().numpy().view(U64), oi)
    np.testing.assert_array_equal(gm.cpu().numpy().view(np.uint32), om)
    corpus.close()
