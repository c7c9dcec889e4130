"""CPU restatement of the ARITHMETIC of the tensor-core Hamming filter (ucfp_b200/csrc/hamming.cu, hamming_mma_scan_kernel):
the +-1 / packed operand encoding, the 16-bit accumulator image, the two s16 bound tests of the hot loop and the decode of
the cold path.  It pins, without a GPU, the claims the kernel's exactness rests on:

  * D = -x_a + 64 x_b with x = 64 - 2 dist, |D| <= 4160, element values +-63 / +-65 fit s8;
  * the hot test has NO false negatives for any thr (and false positives only where documented: x_a = -64);
  * the cold-path decode returns both true distances whenever u != 0, and u == 0 happens only for dist_a in {0, 64}.
"""
import numpy as np

U64 = np.uint64


def popcount(x: np.ndarray) -> np.ndarray:
    x = x.astype(U64)
    c = np.zeros(x.shape, dtype=np.int64)
    for i in range(64):
        c += ((x >> U64(i)) & U64(1)).astype(np.int64)
    return c


def bits_pm1(code: np.ndarray) -> np.ndarray:
    """bit set -> +1, clear -> -1, shape (..., 64)"""
    b = ((code[..., None] >> np.arange(64, dtype=U64)) & U64(1)).astype(np.int64)
    return 2 * b - 1


def operand_row(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """element k = -a_k + 64 b_k, built the way mma_pack4 builds it: 0xC1 ^ (abit * 0x7E) ^ (bbit * 0x80), as s8"""
    abit = ((a[..., None] >> np.arange(64, dtype=U64)) & U64(1)).astype(np.uint8)
    bbit = ((b[..., None] >> np.arange(64, dtype=U64)) & U64(1)).astype(np.uint8)
    byte = np.uint8(0xC1) ^ (abit * np.uint8(0x7E)) ^ (bbit * np.uint8(0x80))
    return byte.view(np.int8).astype(np.int64)


def s16(v: np.ndarray) -> np.ndarray:
    return ((v.astype(np.int64) + 0x8000) % 0x10000) - 0x8000


def bounds(thr_hot):
    """hi16 / lo16 of the kernel prologue for one hot-test bound (None = never)"""
    if thr_hot is None:
        return 0x8000, -0x8001     # stored as 0x7FFF / -0x8000: no halfword exceeds the one or falls below the other
    if thr_hot >= 64:
        return -0x7FFF, 0
    tau = 64 - 2 * thr_hot
    return 64 * (tau - 1), (-tau * 512) | 0x1FF


def make_cases(rng, n):
    q = rng.integers(0, 2**63, dtype=np.int64).astype(U64) | (U64(rng.integers(0, 2)) << U64(63))
    a = rng.integers(0, 2**63, size=n, dtype=np.int64).astype(U64) ^ (rng.integers(0, 2, size=n).astype(U64) << U64(63))
    b = rng.integers(0, 2**63, size=n, dtype=np.int64).astype(U64) ^ (rng.integers(0, 2, size=n).astype(U64) << U64(63))
    # near and far neighbours of q at every distance, on both sides of the pair
    for d in range(65):
        mask = U64(0)
        for bit in rng.choice(64, d, replace=False):
            mask |= U64(1) << U64(int(bit))
        a[d] = q ^ mask
        b[64 + d] = q ^ mask
        a[130 + d] = q ^ mask
        b[130 + d] = q ^ U64(2**64 - 1) ^ mask
    return q, a, b


def test_operand_encoding_and_accumulator_range():
    rng = np.random.default_rng(1)
    q, a, b = make_cases(rng, 4096)
    row = operand_row(a, b)
    assert set(np.unique(row)) <= {-65, -63, 63, 65}
    np.testing.assert_array_equal(row, -bits_pm1(a) + 64 * bits_pm1(b))
    D = (row * bits_pm1(np.array([q]))).sum(axis=-1)
    xa, xb = 64 - 2 * popcount(a ^ q), 64 - 2 * popcount(b ^ q)
    np.testing.assert_array_equal(D, -xa + 64 * xb)
    assert np.abs(D).max() <= 4160


def test_hot_test_has_no_false_negatives_and_only_documented_false_positives():
    rng = np.random.default_rng(2)
    q, a, b = make_cases(rng, 1 << 15)
    da, db = popcount(a ^ q), popcount(b ^ q)
    D = -(64 - 2 * da) + 64 * (64 - 2 * db)
    if len(D) % 2:
        D = D[:-1]
    # register image: odd column in the upper halfword, even column in the lower one
    lo_half, hi_half = D[0::2] & 0xFFFF, D[1::2] & 0xFFFF
    reg = (hi_half << 16) | lo_half
    y_lo, y_hi = s16(reg & 0xFFFF), s16(reg >> 16)
    sh = (reg * 512) & 0xFFFFFFFF
    x_lo, x_hi = s16(sh & 0xFFFF), s16(sh >> 16)
    np.testing.assert_array_equal(y_lo, D[0::2])
    np.testing.assert_array_equal(y_hi, D[1::2])
    for thr in list(range(0, 66)) + [None]:
        hi16, lo16 = bounds(thr)
        fired_lo = (y_lo >= hi16) | (x_lo <= lo16)       # tests of the even column of each register
        fired_hi = (y_hi >= hi16) | (x_hi <= lo16)
        if thr is None:
            want = np.zeros(len(D), dtype=bool)
        else:
            want = (da[: len(D)] <= thr) | (db[: len(D)] <= thr)
        got = np.empty(len(D), dtype=bool)
        got[0::2], got[1::2] = fired_lo, fired_hi
        assert not (want & ~got).any(), f"false negative at thr {thr}"
        extra = got & ~want
        if thr is None:      # never means never: also for x_a = +-64, whose low field is the most negative halfword there is
            assert not extra.any()
        else:                # false positives: dist_a == 64 (aliases dist 0); the y test may also admit dist_b <= thr when x_a = -64
            assert ((da[: len(D)][extra] == 64) | (db[: len(D)][extra] <= thr + 1)).all(), thr
            assert (da[: len(D)][extra] == 64).all() or thr >= 31


def test_cold_path_decode_is_exact():
    rng = np.random.default_rng(3)
    q, a, b = make_cases(rng, 1 << 15)
    da, db = popcount(a ^ q), popcount(b ^ q)
    D = -(64 - 2 * da) + 64 * (64 - 2 * db)
    u = (D ^ 64) & 127
    ambiguous = u == 0
    assert set(np.unique(da[ambiguous])) <= {0, 64}
    assert not ambiguous[(da != 0) & (da != 64)].any()
    xa = 64 - u
    xb = (D + xa) >> 6
    np.testing.assert_array_equal((u >> 1)[~ambiguous], da[~ambiguous])
    np.testing.assert_array_equal(((64 - xb) >> 1)[~ambiguous], db[~ambiguous])


def test_spread4_bit_trick():
    """(nibble * 0x00204081) & 0x01010101 puts bit i of the nibble into byte i"""
    for nib in range(16):
        t = (nib * 0x00204081) & 0x01010101
        assert [(t >> (8 * i)) & 0xFF for i in range(4)] == [(nib >> i) & 1 for i in range(4)]
        w = ~(t * 0xFE) & 0xFFFFFFFF                     # query rows: set -> 0x01, clear -> 0xFF
        assert [(w >> (8 * i)) & 0xFF for i in range(4)] == [1 if (nib >> i) & 1 else 0xFF for i in range(4)]
