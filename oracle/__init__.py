"""ctypes loader for the CPU oracle (oracle/ucfp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by
bench.py's cpu_baseline / --impl reference legs.  The product package
(ucfp_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libucfp_oracle.so")
_lib = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc only)."""
    src = os.path.join(_HERE, "ucfp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.ucfp_oracle_splitmix64.restype = C.c_uint64
        L.ucfp_oracle_splitmix64.argtypes = [C.c_uint64, C.c_uint64]
        L.ucfp_oracle_fill_u64.argtypes = [u64p, C.c_size_t, C.c_uint64, C.c_uint64]
        L.ucfp_oracle_fill_u64_mt.argtypes = [u64p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int]
        for name in ("ucfp_oracle_hamming_topk", "ucfp_oracle_jaccard_topk"):
            getattr(L, name).argtypes = [u64p, u64p, C.c_uint64, C.c_size_t, u64p, C.c_size_t, C.c_size_t,
                                         u64p, u32p, C.c_int]
            getattr(L, name).restype = None
        L.ucfp_oracle_multihash_score.restype = C.c_float
        L.ucfp_oracle_multihash_score.argtypes = [u64p, u64p, C.POINTER(C.c_float), C.c_uint32]
        L.ucfp_oracle_dot_product.restype = C.c_float
        L.ucfp_oracle_dot_product.argtypes = [f32p, f32p, C.c_size_t]
        L.ucfp_oracle_l2_norm.restype = C.c_float
        L.ucfp_oracle_l2_norm.argtypes = [f32p, C.c_size_t]
        L.ucfp_oracle_cosine_topk.argtypes = [f32p, u64p, C.c_uint64, C.c_size_t, C.c_size_t, f32p, C.c_size_t,
                                              C.c_size_t, C.c_int, u64p, f32p, u32p, C.c_int]
        L.ucfp_oracle_cosine_topk.restype = None
        L.ucfp_oracle_gray.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p]
        L.ucfp_oracle_triangle_taps.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), f32p, C.c_int]
        L.ucfp_oracle_triangle_taps.restype = C.c_int
        L.ucfp_oracle_resize_triangle.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, u8p]
        L.ucfp_oracle_ahash_bits.restype = C.c_uint64
        L.ucfp_oracle_ahash_bits.argtypes = [u8p]
        L.ucfp_oracle_dhash_bits.restype = C.c_uint64
        L.ucfp_oracle_dhash_bits.argtypes = [u8p]
        L.ucfp_oracle_phash_bits.restype = C.c_uint64
        L.ucfp_oracle_phash_bits.argtypes = [u8p, f32p]
        L.ucfp_oracle_image_multihash.restype = C.c_int
        L.ucfp_oracle_image_multihash.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u64p]
        L.ucfp_oracle_image_multihash_batch.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                                        u64p, C.c_int]
        L.ucfp_oracle_image_multihash_batch.restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ---------------------------------------------------------------- PRNG -------
def splitmix64(seed: int, index: int) -> int:
    return int(lib().ucfp_oracle_splitmix64(seed & (2**64 - 1), index & (2**64 - 1)))


def fill_u64(n: int, seed: int, start: int = 0, threads: int = 0, out: np.ndarray = None) -> np.ndarray:
    """out[i] = splitmix64(seed, start + i); large fills use all host threads (threads=0: decide by size).
    `out`: a C-contiguous uint64 array of >= n elements to fill in place (its first n elements are returned)."""
    if out is None:
        out = np.empty(n, dtype=np.uint64)
    else:
        assert out.dtype == np.uint64 and out.flags["C_CONTIGUOUS"] and out.size >= n
        out = out.reshape(-1)[:n]
    if threads == 0:
        threads = host_threads() if n >= (1 << 22) else 1
    if threads > 1:
        lib().ucfp_oracle_fill_u64_mt(_p(out, u64p), n, seed, start, threads)
    else:
        lib().ucfp_oracle_fill_u64(_p(out, u64p), n, seed, start)
    return out


# ---------------------------------------------------------------- scans ------
def _ids_arg(ids):
    if ids is None:
        return None, None
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    return ids, _p(ids, u64p)


def hamming_topk(codes, queries, k, ids=None, id_base=0, threads=1):
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    queries = np.ascontiguousarray(queries, dtype=np.uint64)
    nq = queries.shape[0]
    ids_out = np.full((nq, k), 2**64 - 1, dtype=np.uint64)
    dist_out = np.full((nq, k), 2**32 - 1, dtype=np.uint32)
    keep, idp = _ids_arg(ids)
    lib().ucfp_oracle_hamming_topk(_p(codes, u64p), idp, id_base, codes.shape[0], _p(queries, u64p), nq, k,
                                   _p(ids_out, u64p), _p(dist_out, u32p), threads)
    return ids_out, dist_out


def jaccard_topk(sigs, queries, k, ids=None, id_base=0, threads=1):
    sigs = np.ascontiguousarray(sigs, dtype=np.uint64).reshape(-1, 128)
    queries = np.ascontiguousarray(queries, dtype=np.uint64).reshape(-1, 128)
    nq = queries.shape[0]
    ids_out = np.full((nq, k), 2**64 - 1, dtype=np.uint64)
    m_out = np.full((nq, k), 2**32 - 1, dtype=np.uint32)
    keep, idp = _ids_arg(ids)
    lib().ucfp_oracle_jaccard_topk(_p(sigs, u64p), idp, id_base, sigs.shape[0], _p(queries, u64p), nq, k,
                                   _p(ids_out, u64p), _p(m_out, u32p), threads)
    return ids_out, m_out


def dot_product(a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return float(lib().ucfp_oracle_dot_product(_p(a, f32p), _p(b, f32p), a.shape[0]))


def l2_norm(a) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return float(lib().ucfp_oracle_l2_norm(_p(a, f32p), a.shape[0]))


def cosine_topk(rows, queries, k, ids=None, id_base=0, mode=1, threads=1):
    """mode 0 = reference insert_topk tie behaviour, mode 1 = total order (score desc, id asc).
    Returns (ids[nq,k], scores[nq,k], counts[nq])."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    if rows.ndim == 1:
        rows = rows.reshape(0, queries.shape[1] if queries.ndim == 2 else 0)
    n, dim = rows.shape
    queries = queries.reshape(-1, dim) if dim else queries.reshape(queries.shape[0], 0)
    nq = queries.shape[0]
    ids_out = np.full((nq, max(k, 1)), 2**64 - 1, dtype=np.uint64)
    sc_out = np.full((nq, max(k, 1)), -np.inf, dtype=np.float32)
    cnt = np.zeros(nq, dtype=np.uint32)
    keep, idp = _ids_arg(ids)
    lib().ucfp_oracle_cosine_topk(_p(rows, f32p), idp, id_base, n, dim, _p(queries, f32p), nq, k, mode,
                                  _p(ids_out, u64p), _p(sc_out, f32p), _p(cnt, u32p), threads)
    return ids_out[:, :k], sc_out[:, :k], cnt


# ---------------------------------------------------------------- images -----
def gray(rgb: np.ndarray) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    out = np.empty((h, w), dtype=np.uint8)
    lib().ucfp_oracle_gray(_p(rgb, u8p), w, h, 3 * w, _p(out, u8p))
    return out


def triangle_taps(src: int, dst: int, o: int):
    w = np.zeros(4 * src + 8, dtype=np.float32)
    left = C.c_int(0)
    n = lib().ucfp_oracle_triangle_taps(src, dst, o, C.byref(left), _p(w, f32p), w.shape[0])
    return left.value, w[:n].copy()


def resize_triangle(g: np.ndarray, nw: int, nh: int) -> np.ndarray:
    g = np.ascontiguousarray(g, dtype=np.uint8)
    h, w = g.shape
    out = np.empty((nh, nw), dtype=np.uint8)
    lib().ucfp_oracle_resize_triangle(_p(g, u8p), w, h, w, nw, nh, _p(out, u8p))
    return out


def ahash_bits(g8) -> int:
    g8 = np.ascontiguousarray(g8, dtype=np.uint8).reshape(64)
    return int(lib().ucfp_oracle_ahash_bits(_p(g8, u8p)))


def dhash_bits(g98) -> int:
    g98 = np.ascontiguousarray(g98, dtype=np.uint8).reshape(72)
    return int(lib().ucfp_oracle_dhash_bits(_p(g98, u8p)))


def phash_bits(g32, want_coeff=False):
    g32 = np.ascontiguousarray(g32, dtype=np.uint8).reshape(1024)
    co = np.zeros(64, dtype=np.float32)
    bits = int(lib().ucfp_oracle_phash_bits(_p(g32, u8p), _p(co, f32p)))
    return (bits, co.reshape(8, 8)) if want_coeff else bits


def image_multihash(rgb: np.ndarray) -> np.ndarray:
    """rgb: (h, w, 3) u8 -> 51 u64: ahash[17] | phash[17] | dhash[17]."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    out = np.zeros(51, dtype=np.uint64)
    rc = lib().ucfp_oracle_image_multihash(_p(rgb, u8p), w, h, 3 * w, _p(out, u64p))
    if rc != 0:
        raise ValueError("image too small for a 4x4 block grid")
    return out


def image_multihash_batch(rgb: np.ndarray, threads: int = 1) -> np.ndarray:
    """rgb: (n, h, w, 3) u8 -> (n, 51) u64."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    n, h, w, _ = rgb.shape
    out = np.zeros((n, 51), dtype=np.uint64)
    lib().ucfp_oracle_image_multihash_batch(_p(rgb, u8p), n, w, h, 3 * w, 3 * w * h, _p(out, u64p), threads)
    return out


# ---------------------------------------------------------------- multi-hash compare / re-rank (spec section 10) ------
MULTIHASH_DEFAULTS = {"ahash_weight": 0.1, "phash_weight": 0.4, "dhash_weight": 0.3, "global_weight": 0.1, "block_weight": 0.1,
                      "block_distance_threshold": 12}   # web/src/lib/docs/api-reference-image.md:51-62


def _mh_cfg(cfg):
    c = dict(MULTIHASH_DEFAULTS)
    c.update(cfg or {})
    arr = (C.c_float * 5)(c["ahash_weight"], c["phash_weight"], c["dhash_weight"], c["global_weight"], c["block_weight"])
    return arr, int(c["block_distance_threshold"])


def multihash_score(x: np.ndarray, y: np.ndarray, cfg=None) -> float:
    """x, y: 51 u64 each (ahash[17] | phash[17] | dhash[17])."""
    arr, thr = _mh_cfg(cfg)
    x = np.ascontiguousarray(x, dtype=np.uint64).reshape(51)
    y = np.ascontiguousarray(y, dtype=np.uint64).reshape(51)
    return float(lib().ucfp_oracle_multihash_score(_p(x, u64p), _p(y, u64p), arr, thr))


def multihash_rerank(bundles: np.ndarray, queries: np.ndarray, k_prime: int, k: int, ids=None, cfg=None, threads: int = 1):
    """Spec section 10: candidates = Hamming top-k' on the PHash global hash (word 17), re-ranked by the blended score;
    order (score desc, id asc).  -> (ids [nq, k] u64, scores [nq, k] f32)."""
    bundles = np.ascontiguousarray(bundles, dtype=np.uint64).reshape(-1, 51)
    queries = np.ascontiguousarray(queries, dtype=np.uint64).reshape(-1, 51)
    n, nq = len(bundles), len(queries)
    rid = np.arange(n, dtype=np.uint64) if ids is None else np.ascontiguousarray(ids, dtype=np.uint64)
    # coarse pass: ranked by (distance, RECORD id) -- the candidate set must not depend on the row order
    cand_ids, _ = hamming_topk(np.ascontiguousarray(bundles[:, 17]), np.ascontiguousarray(queries[:, 17]), k_prime, ids=rid, threads=threads)
    row_of = {int(i): r for r, i in enumerate(rid)}
    rows = np.array([[row_of.get(int(i), 2**64 - 1) for i in cq] for cq in cand_ids], dtype=np.uint64).reshape(nq, -1)
    out_i = np.full((nq, k), np.uint64(2**64 - 1), dtype=np.uint64)
    out_s = np.full((nq, k), -np.inf, dtype=np.float32)
    for q in range(nq):
        cand = [int(r) for r in rows[q] if r != np.uint64(2**64 - 1)]
        scored = sorted(((-np.float32(multihash_score(bundles[r], queries[q], cfg)), int(rid[r])) for r in cand))[:k]
        for j, (ns, i) in enumerate(scored):
            out_i[q, j], out_s[q, j] = i, -ns
    return out_i, out_s
