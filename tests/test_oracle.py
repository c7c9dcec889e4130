"""CPU-only checks of the oracle (oracle/ucfp_oracle.c): the reference's own known-answer tests for the path
whose arithmetic is in-tree (cosine), independent cross-checks of the restated published algorithms, and the
committed golden vectors.  Runs without a GPU."""
import json
import os

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))


def ramp(w, h):
    y, x = np.mgrid[0:h, 0:w]
    return np.stack([x % 256, y % 256, np.full_like(x, 128)], -1).astype(np.uint8)


def noise(w, h, seed):
    return oracle.fill_u64((w * h * 3 + 7) // 8, seed).view(np.uint8)[: w * h * 3].reshape(h, w, 3).copy()


# ---- cosine: pinned by the reference's tests ---------------------------------------------------------------
def test_ref_upsert_and_knn_round_trip():
    """reference src/index/embedded/mod.rs:523-544."""
    rows = np.array([[1, 0, 0], [0, 1, 0], [0.7, 0.7, 0]], np.float32)
    ids, sc, cnt = oracle.cosine_topk(rows, np.array([[0.6, 0.6, 0.0]], np.float32), 2, ids=[100, 200, 300], mode=0)
    assert cnt[0] == 2 and ids[0, 0] == 300 and sc[0, 0] > sc[0, 1]


def test_ref_knn_single_tenant_and_delete_shapes():
    """reference embedded/mod.rs:547-573: one matching record of dimension 2 -> exactly one hit."""
    ids, sc, cnt = oracle.cosine_topk(np.array([[1.0, 0.0]], np.float32), np.array([[1.0, 0.0]], np.float32), 10, ids=[1], mode=0)
    assert cnt[0] == 1 and ids[0, 0] == 1 and (ids[0, 1:] == np.uint64(2**64 - 1)).all()
    ids, sc, cnt = oracle.cosine_topk(np.array([[0.0, 1.0]], np.float32), np.array([[1.0, 0.0]], np.float32), 10, ids=[2], mode=0)
    assert cnt[0] == 1 and ids[0, 0] == 2 and sc[0, 0] == 0.0


def test_ref_server_query_top_hit():
    """reference src/server/tests.rs:53-113: the record aligned with the query is the top hit."""
    rows = np.eye(4, dtype=np.float32)[:3]
    ids, _, _ = oracle.cosine_topk(rows, np.array([[0.1, 0.9, 0, 0]], np.float32), 10, ids=[100, 200, 300], mode=0)
    assert ids[0, 0] == 200


def test_ref_empty_and_zero_queries_return_nothing():
    """embedded/mod.rs:275 (k == 0), :284 (zero-norm query), :328-330 (zero-norm rows skipped)."""
    rows = np.array([[1, 0], [0, 0], [0, 1]], np.float32)
    _, _, cnt = oracle.cosine_topk(rows, np.array([[0.0, 0.0]], np.float32), 5, mode=0)
    assert cnt[0] == 0
    ids, _, cnt = oracle.cosine_topk(rows, np.array([[1.0, 1.0]], np.float32), 5, mode=0)
    assert cnt[0] == 2 and 1 not in ids[0, :2]
    ids, _, cnt = oracle.cosine_topk(rows, np.array([[1.0, 1.0]], np.float32), 0, mode=0)
    assert ids.shape == (1, 0)


def test_ref_insert_topk_tie_behaviour():
    """SURVEY A7 / embedded/mod.rs:484-495: new ties go in front of old ties, a full buffer admits only strictly
    better scores, so membership among boundary ties depends on scan order: sequential -> [5,2,4,3], a 2-way
    fold/reduce -> [5,2,3,1].  Mode 1 (what the GPU guarantees) is (score desc, id asc) -> [2,5,1,3]."""
    a, b = np.array([1.0, 1.0], np.float32), np.array([1.0, 0.2], np.float32)
    rows = np.stack([a, b, a, a, b, a])            # ids 1..6: scores s(a) < s(b) for the query below
    q = np.array([[1.0, 0.0]], np.float32)
    ids = [1, 2, 3, 4, 5, 6]
    seq, _, _ = oracle.cosine_topk(rows, q, 4, ids=ids, mode=0, threads=1)
    assert seq[0].tolist() == [5, 2, 4, 3]
    par, _, _ = oracle.cosine_topk(rows, q, 4, ids=ids, mode=0, threads=2)
    assert par[0].tolist() == [5, 2, 3, 1]
    tot, _, _ = oracle.cosine_topk(rows, q, 4, ids=ids, mode=1, threads=2)
    assert tot[0].tolist() == [2, 5, 1, 3]


def test_dot_product_is_eight_lane_order():
    """embedded/mod.rs:454-472 against an independent numpy restatement, including a remainder."""
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 512, 515):
        a, b = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
        accs = np.zeros(8, np.float32)
        for c in range(n // 8):
            accs = (accs + a[c * 8:c * 8 + 8] * b[c * 8:c * 8 + 8]).astype(np.float32)
        s = np.float32(0)
        for j in range(8):
            s = np.float32(s + accs[j])
        for i in range(n // 8 * 8, n):
            s = np.float32(s + np.float32(a[i] * b[i]))
        assert oracle.dot_product(a, b) == float(s)
        assert oracle.l2_norm(a) == float(np.sqrt(np.float32(oracle.dot_product(a, a)), dtype=np.float32))


# ---- MinHash layout: pinned by the reference's golden bytes ----------------------------------------------
def test_minhash_sig_layout_from_reference_golden():
    """server/tests.rs:1153-1162: 1032 bytes, hex[..32] == 0100000000000000 a26accc88c8a8106 -> schema u16 = 1,
    6 pad bytes, then little-endian u64 slots."""
    from ucfp_b200.image import minhash_payload_of
    from ucfp_b200.core import Error
    blob = bytes.fromhex("0100000000000000a26accc88c8a8106") + bytes(1032 - 16)
    slots = minhash_payload_of(blob)
    assert slots.shape == (128,) and slots[0] == 0x06818A8CC8CC6AA2 and slots[1] == 0
    with pytest.raises(Error):
        minhash_payload_of(blob[:-1])
    with pytest.raises(Error):
        minhash_payload_of(b"\x02" + blob[1:])


# ---- image pipeline: independent cross-checks of the restated published algorithms -----------------------
def test_luma_matches_image_crate_formula():
    img = noise(64, 48, 5)
    i = img.astype(np.int64)
    want = ((2126 * i[..., 0] + 7152 * i[..., 1] + 722 * i[..., 2]) // 10000).astype(np.uint8)
    np.testing.assert_array_equal(oracle.gray(img), want)


@pytest.mark.parametrize("src,dst", [(256, 32), (1024, 32), (1024, 8), (256, 9), (100, 32), (37, 32), (16, 32), (9, 8), (33, 32), (4, 8)])
def test_triangle_taps_match_float64_restatement(src, dst):
    """Weights of `image` 0.25's Triangle sampler: normalised, windowed as published, close to the f64 formula."""
    ratio = src / dst
    sratio = max(ratio, 1.0)
    for o in range(dst):
        left, w = oracle.triangle_taps(src, dst, o)
        c = (o + 0.5) * ratio
        lo = min(max(int(np.floor(c - sratio)), 0), src - 1)
        hi = min(max(int(np.ceil(c + sratio)), lo + 1), src)
        # f32 vs f64 may disagree on a window edge by one tap of (near-)zero weight
        assert abs(left - lo) <= 1 and abs((left + len(w)) - hi) <= 1
        x = (np.arange(left, left + len(w)) - (c - 0.5)) / sratio
        ref = np.clip(1 - np.abs(x), 0, None)
        ref /= ref.sum()
        np.testing.assert_allclose(w, ref, atol=2e-6)
        assert abs(float(w.sum()) - 1.0) < 1e-5


def test_resize_close_to_pillow_and_float64():
    from PIL import Image
    g = oracle.gray(noise(256, 192, 9))
    for nw, nh in ((32, 32), (9, 8), (8, 8)):
        got = oracle.resize_triangle(g, nw, nh).astype(int)
        pil = np.asarray(Image.fromarray(g).resize((nw, nh), Image.BILINEAR)).astype(int)
        assert np.abs(got - pil).max() <= 1          # Pillow: same Triangle support, fixed-point, other pass order
    same = oracle.resize_triangle(g, 256, 192)
    np.testing.assert_array_equal(same, g)          # imageops::resize copies when the size is unchanged


def test_phash_dct_matches_scipy_and_median_rule():
    import scipy.fft
    g32 = oracle.resize_triangle(oracle.gray(noise(256, 256, 3)), 32, 32)
    bits, co = oracle.phash_bits(g32, want_coeff=True)
    ref = scipy.fft.dctn(g32.astype(np.float64), type=2)[:8, :8] / 4.0   # un-normalised DCT-II
    np.testing.assert_allclose(co, ref, rtol=1e-5, atol=1e-2)
    s = np.sort(co.reshape(-1))
    med = np.float32((s[31] + s[32]) * np.float32(0.5))
    want = sum(1 << i for i, v in enumerate(co.reshape(-1)) if v > med)
    assert bits == want and bin(bits).count("1") <= 32


def test_ahash_dhash_definitions():
    rng = np.random.default_rng(1)
    g8 = rng.integers(0, 256, 64, dtype=np.uint8)
    mean_f = g8.astype(np.float32).sum() / np.float32(64)
    assert oracle.ahash_bits(g8) == sum(1 << i for i, p in enumerate(g8) if np.float32(p) > mean_f)
    assert oracle.ahash_bits(g8) == sum(1 << i for i, p in enumerate(g8) if int(p) > int(g8.astype(int).sum()) // 64)  # image.rs:317
    g98 = rng.integers(0, 256, 72, dtype=np.uint8).reshape(8, 9)
    assert oracle.dhash_bits(g98) == sum(1 << (8 * r + c) for r in range(8) for c in range(8) if g98[r, c] > g98[r, c + 1])
    assert oracle.dhash_bits(np.tile(np.arange(9, dtype=np.uint8), (8, 1))) == 0          # the reference ramp: never left > right


def test_bundle_layout_and_reference_size_checks():
    """server/tests.rs:1203-1207: 536-byte bundle, hex length 1072, 0 < ahash_mean < 256; algorithmView.ts:7-16."""
    from ucfp_b200 import image as gi
    words = oracle.image_multihash(ramp(256, 256))
    blob = gi.pack_multihash(bytes(range(32)), words)
    assert len(blob) == 536 and len(blob.hex()) == 1072
    assert blob[:32] == bytes(range(32)) and blob[32:64] == bytes(range(32))     # each ImageFingerprint repeats `exact`
    assert int.from_bytes(blob[64:72], "little") == int(words[0])                # ahash.global_hash @ 32 + 32
    assert gi.global_hash_of(blob, gi.ALGORITHM_MULTIHASH) == int(words[17])     # PHash global @ 232
    single = gi.pack_image_fingerprint(bytes(32), words[34:51])
    assert len(single) == 168 and gi.global_hash_of(single, gi.ALGORITHM_DHASH) == int(words[34])
    g8 = oracle.resize_triangle(oracle.gray(ramp(256, 256)), 8, 8)
    assert 0 < int(g8.astype(int).sum()) // 64 < 256


# ---- committed golden vectors ------------------------------------------------------------------------------
def test_golden_prng_and_images():
    assert [f"{int(v):016x}" for v in oracle.fill_u64(8, GOLD["prng"]["seed"])] == GOLD["prng"]["first8"]
    makers = {"ramp256": lambda: ramp(256, 256), "ramp64": lambda: ramp(64, 64), "ramp300x200": lambda: ramp(300, 200),
              "noise256_s1": lambda: noise(256, 256, 1), "noise1024_s2": lambda: noise(1024, 1024, 2),
              "noise37x53_s3": lambda: noise(37, 53, 3), "noise640x480_s4": lambda: noise(640, 480, 4)}
    for e in GOLD["images"]:
        got = [f"{int(v):016x}" for v in oracle.image_multihash(makers[e["name"]]())]
        assert got == e["words"], e["name"]


def test_golden_scans():
    h = GOLD["hamming"]
    ids, d = oracle.hamming_topk(oracle.fill_u64(h["n"], h["seed_codes"]), oracle.fill_u64(4, h["seed_queries"]), h["k"], threads=3)
    assert ids.tolist() == h["ids"] and d.tolist() == h["dist"]
    j = GOLD["jaccard"]
    sig = oracle.fill_u64(j["n"] * 128, j["seed_sigs"]).reshape(j["n"], 128)
    qs = oracle.fill_u64(2 * 128, j["seed_queries"]).reshape(2, 128)
    sig[1234, :100] = qs[0, :100]
    sig[77, 5:70] = qs[1, 5:70]
    sig[78, 5:70] = qs[1, 5:70]
    ids, m = oracle.jaccard_topk(sig, qs, j["k"], threads=2)
    assert ids.tolist() == j["ids"] and m.tolist() == j["matches"]
    assert m[0, 0] == 100 and ids[1, :2].tolist() == [77, 78]                   # equal matches -> id ascending


def test_thread_count_does_not_change_total_order_results():
    codes, q = oracle.fill_u64(30_000, 3) & np.uint64(0xFFFF), oracle.fill_u64(5, 4) & np.uint64(0xFFFF)   # heavy ties
    base = oracle.hamming_topk(codes, q, 20, threads=1)
    for t in (2, 5, 8):
        got = oracle.hamming_topk(codes, q, 20, threads=t)
        np.testing.assert_array_equal(got[0], base[0])
        np.testing.assert_array_equal(got[1], base[1])
