"""Record-range sharding of a scan corpus over the GPUs of one box and the host-side reference
of the merge step, used by the CPU (gloo) tests of the N > 1 protocol.

Rank r of G holds rows [r*N/G, (r+1)*N/G) and reports GLOBAL record ids; every rank receives the full query
batch, scans its slice, all-gathers the per-rank top-k (Q x k x (u64 id, u32/f32 key)) and runs the same
deterministic merge, so any rank can answer and the result is byte-identical to the 1-GPU result."""
from __future__ import annotations

from typing import Tuple

import numpy as np

ID_NONE = np.uint64(2**64 - 1)


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    return n_total * rank // world, n_total * (rank + 1) // world


def merge_topk_host(ids: np.ndarray, keys: np.ndarray, k: int, descending: bool) -> Tuple[np.ndarray, np.ndarray]:
    """ids, keys: [parts, nq, k] as gathered.  Returns [nq, k] under (key best-first, id asc), sentinels last --
    the contract of ucfp_merge_topk_u32 / _f32."""
    parts, nq, kk = ids.shape
    out_i = np.full((nq, k), ID_NONE, dtype=np.uint64)
    sentinel = np.float32(-np.inf) if keys.dtype == np.float32 else np.uint32(2**32 - 1)
    out_k = np.full((nq, k), sentinel, dtype=keys.dtype)
    for q in range(nq):
        i = ids[:, q, :].reshape(-1)
        v = keys[:, q, :].reshape(-1)
        ok = i != ID_NONE
        i, v = i[ok], v[ok]
        primary = -v.astype(np.float64) if descending else v.astype(np.float64)
        order = np.lexsort((i, primary))[:k]
        out_i[q, : len(order)] = i[order]
        out_k[q, : len(order)] = v[order]
    return out_i, out_k
