"""CPU restatement of the overflow rule of the candidate lists (ucfp_b200/csrc/topk_select.cuh, cand_append): a full list keeps a
RESERVOIR sample (Algorithm R, a hash of the arrival number as the random source) of everything that was admitted, so that the bound
the next compaction derives from it shrinks a flood whatever the order of arrival.  These tests pin the arithmetic of the rule and
the property the re-scan rounds rely on; the kernels themselves are covered by the GPU flood tests."""
import numpy as np

CAP = 4096
LO = CAP >> 2


def _hash(pos: np.ndarray) -> np.ndarray:
    h = (pos.astype(np.uint64) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    return h


def _final_list(order: np.ndarray) -> np.ndarray:
    """Replays cand_append for arrivals 0..len(order)-1 carrying the values `order`; returns the list's final content."""
    lst = np.full(CAP, -1, dtype=np.int64)
    lst[:CAP] = order[:CAP]
    pos = np.arange(CAP, len(order), dtype=np.uint64)
    j = (_hash(pos) * (pos - np.uint64(LO) + np.uint64(1))) >> np.uint64(32)      # __umulhi(h, pos - lo + 1)
    sel = j < np.uint64(CAP - LO)
    slots = (np.uint64(LO) + j[sel]).astype(np.int64)
    vals = order[CAP:][sel]
    lst[slots] = vals          # numpy keeps the last write per slot: arrival order, as on the device up to races between equals
    return lst


def test_first_quarter_is_never_overwritten():
    order = np.arange(200_000, dtype=np.int64)
    lst = _final_list(order)
    assert (lst[:LO] == order[:LO]).all()          # the k kept entries (k <= cap / 4) and the earliest arrivals stay
    assert (lst >= 0).all()


def test_sample_bound_shrinks_a_flood_whatever_the_arrival_order():
    """value = rank of the row (0 = best).  The 10th best of the final list must have a rank of the order of k M / (0.75 cap), far
    below the list's capacity, for: best rows last, best rows first, best rows in the middle (what a two-wave launch produces: the
    case a last-writer-wins rule got wrong by a factor of 40)."""
    m, k = 600_000, 10
    asc = np.arange(m, dtype=np.int64)
    orders = {
        "best last": asc[::-1].copy(),
        "best first": asc.copy(),
        "best in the middle": np.concatenate([asc[300_000:], asc[:110_000], asc[110_000:300_000]]),
        "random": np.random.default_rng(1).permutation(m).astype(np.int64),
    }
    expected = k * m / (0.75 * CAP)
    for name, order in orders.items():
        kth = np.sort(_final_list(order))[k - 1]
        assert kth < CAP, (name, kth)                      # the re-scan under this bound fits the list
        assert kth < 3 * expected, (name, kth, expected)


def test_selection_probability_matches_algorithm_r():
    """Arrival number p >= cap replaces a reservoir slot with probability (cap - lo) / (p - lo + 1)."""
    pos = np.arange(CAP, 2_000_000, dtype=np.uint64)
    j = (_hash(pos) * (pos - np.uint64(LO) + np.uint64(1))) >> np.uint64(32)
    sel = j < np.uint64(CAP - LO)
    want = ((CAP - LO) / (pos.astype(np.float64) - LO + 1)).sum()
    assert abs(sel.sum() - want) < 5 * np.sqrt(want)
    # and the chosen slots are spread over the whole reservoir
    hist = np.bincount((j[sel]).astype(np.int64), minlength=CAP - LO)
    assert hist.min() >= 0 and hist.max() < 40 and (hist > 0).mean() > 0.99
