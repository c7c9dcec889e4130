"""Thin object layer over the C ABI: Context (one per GPU) and Corpus (HBM-resident rows).

Buffers may be numpy arrays (host) or torch CUDA tensors (device); only raw pointers cross the ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _ffi
from ._ffi import check

try:  # torch is plumbing (device memory, streams); the library itself never sees it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _ptr(buf) -> int:
    if buf is None:
        return 0
    if isinstance(buf, np.ndarray):
        if not buf.flags["C_CONTIGUOUS"]:
            raise ValueError("buffer must be C-contiguous")
        return buf.ctypes.data
    if torch is not None and isinstance(buf, torch.Tensor):
        if not buf.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return buf.data_ptr()
    if isinstance(buf, int):
        return buf
    raise TypeError(f"unsupported buffer type {type(buf)}")


def _is_device(buf) -> bool:
    return torch is not None and isinstance(buf, torch.Tensor) and buf.is_cuda


def _jpeg_pixels_upper_bound(blob: np.ndarray) -> int:
    """3 * w * h from the first SOFn marker of a JPEG (0 when none is found)."""
    b = blob.tobytes()
    i = 2
    while i + 9 < len(b) and b[i] == 0xFF:
        m = b[i + 1]
        if m in (0xC0, 0xC1, 0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            return 3 * ((b[i + 5] << 8) | b[i + 6]) * ((b[i + 7] << 8) | b[i + 8])
        if m == 0xD8 or 0xD0 <= m <= 0xD7 or m == 0x01:
            i += 2
            continue
        i += 2 + ((b[i + 2] << 8) | b[i + 3])
    return 0


class Context:
    """ucfp_ctx: binds one CUDA device.  Raises UcfpError(UCFP_E_CUDA) without an sm_100 GPU."""

    def __init__(self, device: int = 0, use_torch_stream: bool = True, _borrowed=None):
        """use_torch_stream=True puts the context in shared-stream mode on torch's current stream (work is then ordered with
        torch ops and device-output calls return asynchronously); False keeps the default pooled mode: per-call lanes with
        their own streams, concurrent calls from several host threads, every call complete on return."""
        self._L = _ffi.lib()
        self._owned = _borrowed is None
        if _borrowed is None:
            h = C.c_void_p()
            check(self._L.ucfp_init(device, C.byref(h)))
        else:
            h = C.c_void_p(_borrowed)
        self._h = h
        self.device = device
        if use_torch_stream and torch is not None and torch.cuda.is_available():
            self.set_stream(torch.cuda.current_stream(device).cuda_stream)

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        """cuda_stream: a cudaStream_t handle (0 = CUDA's default stream); None = the context's own stream."""
        if cuda_stream is None:
            check(self._L.ucfp_ctx_reset_stream(self._h))
        else:
            check(self._L.ucfp_ctx_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self) -> None:
        check(self._L.ucfp_ctx_synchronize(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(self._L.ucfp_ctx_kernel_launches(self._h))

    def last_scan_fallbacks(self) -> int:
        """Queries of the most recent scan whose candidate list overflowed and that were recomputed (re-scan rounds, then exact selection)."""
        n = C.c_uint64(0)
        check(self._L.ucfp_ctx_last_scan_fallbacks(self._h, C.byref(n)))
        return int(n.value)

    def last_scan_exact_selects(self) -> int:
        """Queries of the most recent scan that the re-scan rounds could not settle and the exact multi-pass selection recomputed."""
        n = C.c_uint64(0)
        check(self._L.ucfp_ctx_last_scan_exact_selects(self._h, C.byref(n)))
        return int(n.value)

    def last_scan_stats(self):
        """-> (queries recomputed after an overflow, longest candidate list between two compactions) of the most recent scan."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(self._L.ucfp_ctx_last_scan_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def profile_begin(self) -> None:
        check(self._L.ucfp_ctx_profile_begin(self._h))

    def profile_end(self, kernel_class: int):
        """-> (kernel_ms, algorithmic bytes or flops, launches) of one kernel class since profile_begin."""
        ms, units, n = C.c_double(0), C.c_double(0), C.c_uint64(0)
        check(self._L.ucfp_ctx_profile_end(self._h, kernel_class, C.byref(ms), C.byref(units), C.byref(n)))
        return ms.value, units.value, int(n.value)

    def profile_read(self, kernel_class: int):
        """Like profile_end, without ending the region (several kernel classes of one timed region)."""
        ms, units, n = C.c_double(0), C.c_double(0), C.c_uint64(0)
        check(self._L.ucfp_ctx_profile_read(self._h, kernel_class, C.byref(ms), C.byref(units), C.byref(n)))
        return ms.value, units.value, int(n.value)

    def close(self) -> None:
        if getattr(self, "_h", None):
            if self._owned:
                self._L.ucfp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- image hashing --------------------------------------------------------------------
    def image_hash_uniform(self, pixels, n: int, width: int, height: int, algo_mask: int = _ffi.ALGO_MULTI,
                           row_stride: Optional[int] = None, image_stride: Optional[int] = None, out=None):
        """n equally sized RGB8 images -> (n, 51) u64: ahash[17] | phash[17] | dhash[17]."""
        row_stride = row_stride or 3 * width
        image_stride = image_stride or row_stride * height
        if out is None:
            out = (torch.zeros((n, 51), dtype=torch.int64, device=pixels.device) if _is_device(pixels)
                   else np.zeros((n, 51), dtype=np.uint64))
        check(self._L.ucfp_image_hash_uniform(self._h, _ptr(pixels), n, width, height, row_stride, image_stride,
                                              algo_mask, _ptr(out)))
        return out

    def image_hash_batch(self, images, algo_mask: int = _ffi.ALGO_MULTI) -> Tuple[np.ndarray, np.ndarray]:
        """images: list of (h, w, 3) u8 numpy arrays (any sizes).  Returns ((n, 51) u64, (n,) i32 status)."""
        n = len(images)
        descs = (_ffi.ImageDesc * max(n, 1))()
        keep = []
        for i, im in enumerate(images):
            if im is None:
                descs[i] = _ffi.ImageDesc(None, 0, 0, 0)
                continue
            a = np.ascontiguousarray(im, dtype=np.uint8)
            keep.append(a)
            h, w = (a.shape[0], a.shape[1]) if a.ndim == 3 else (0, 0)
            descs[i] = _ffi.ImageDesc(a.ctypes.data, w, h, 3 * w)
        out = np.zeros((n, 51), dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        check(self._L.ucfp_image_hash_batch(self._h, descs, n, algo_mask, _ptr(out), _ptr(status)))
        return out, status

    def image_hash_jpeg_batch(self, blobs, algo_mask: int = _ffi.ALGO_MULTI, want_pixels: bool = False):
        """Encoded JPEG bytes -> hashes, decoded on the device by nvJPEG.  Returns ((n, 51) u64, (n,) i32 status,
        (n, 2) u32 {w, h}[, list of decoded (h, w, 3) u8 arrays or None])."""
        n = len(blobs)
        keep = [np.frombuffer(b, dtype=np.uint8) for b in blobs]
        ptrs = (C.c_void_p * max(n, 1))(*[k.ctypes.data if len(k) else None for k in keep])
        lens = (C.c_size_t * max(n, 1))(*[len(k) for k in keep])
        out = np.zeros((n, 51), dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        dims = np.zeros((n, 2), dtype=np.uint32)
        if not want_pixels:
            check(self._L.ucfp_image_hash_jpeg_batch(self._h, ptrs, lens, n, algo_mask, _ptr(out), _ptr(status), _ptr(dims), None, 0))
            return out, status, dims
        # a first pass for the dimensions would decode twice; size the pixel buffer from the JPEG headers instead
        cap = 0
        for k in keep:
            cap += _jpeg_pixels_upper_bound(k)
        px = np.zeros(max(cap, 1), dtype=np.uint8)
        check(self._L.ucfp_image_hash_jpeg_batch(self._h, ptrs, lens, n, algo_mask, _ptr(out), _ptr(status), _ptr(dims), _ptr(px), cap))
        images, at = [], 0
        for i in range(n):
            if status[i] != 0:
                images.append(None)
                continue
            w, h = int(dims[i, 0]), int(dims[i, 1])
            images.append(px[at: at + 3 * w * h].reshape(h, w, 3).copy())
            at += 3 * w * h
        return out, status, dims, images

    def multihash_compare(self, a, b, cfg=None):
        """Blended global + block similarity of bundle pairs (docs/HASH_SPEC.md section 10).  a, b: (n, 51) u64."""
        n = a.shape[0]
        out = (torch.empty(n, dtype=torch.float32, device=a.device) if _is_device(a) else np.empty(n, dtype=np.float32))
        c = _ffi.MultiHashConfig.of(cfg)
        check(self._L.ucfp_multihash_compare(self._h, _ptr(a), _ptr(b), n, C.byref(c), _ptr(out)))
        return out

    # ---- merges ---------------------------------------------------------------------------
    def merge_topk_u32(self, ids_in, keys_in, parts: int, nq: int, k: int, descending: bool, ids_out, keys_out):
        check(self._L.ucfp_merge_topk_u32(self._h, _ptr(ids_in), _ptr(keys_in), parts, nq, k, int(descending),
                                          _ptr(ids_out), _ptr(keys_out)))

    def merge_topk_f32(self, ids_in, scores_in, parts: int, nq: int, k: int, ids_out, scores_out):
        check(self._L.ucfp_merge_topk_f32(self._h, _ptr(ids_in), _ptr(scores_in), parts, nq, k, _ptr(ids_out),
                                          _ptr(scores_out)))


class Corpus:
    """ucfp_corpus: rows of one (tenant, kind, dim) resident in HBM."""

    def __init__(self, ctx: Context, kind: int, capacity: int, dim: int = 0):
        self.ctx, self.kind, self.dim, self.capacity = ctx, kind, dim, capacity
        self._L = ctx._L
        h = C.c_void_p()
        check(self._L.ucfp_corpus_create(ctx._h, kind, dim, capacity, C.byref(h)))
        self._h = h

    def __len__(self) -> int:
        return int(self._L.ucfp_corpus_size(self._h))

    def append(self, rows, ids=None) -> None:
        n = rows.shape[0]
        check(self._L.ucfp_corpus_append(self._h, _ptr(ids), _ptr(rows), n))

    def append_strided(self, records, record_stride: int, field_offset: int, n: int, ids=None) -> None:
        """Bulk hydration: row i = the row-sized field at `field_offset` of record i (records `record_stride` apart)."""
        check(self._L.ucfp_corpus_append_strided(self._h, _ptr(ids), _ptr(records), record_stride, field_offset, n))

    def append_synthetic(self, seed: int, start_row: int, n: int) -> None:
        check(self._L.ucfp_corpus_append_synthetic(self._h, seed, start_row, n))

    def upsert(self, ids, rows) -> int:
        """Insert-or-replace by record id (IndexBackend::upsert).  Returns the number of rows replaced in place."""
        n = len(ids)
        rep = C.c_uint64(0)
        check(self._L.ucfp_corpus_upsert(self._h, _ptr(ids), _ptr(rows), n, C.byref(rep)))
        return int(rep.value)

    def delete(self, ids) -> int:
        """Removes rows by record id (IndexBackend::delete, idempotent).  Returns the number of rows removed."""
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        rem = C.c_uint64(0)
        check(self._L.ucfp_corpus_delete(self._h, _ptr(ids), len(ids), C.byref(rem)))
        return int(rem.value)

    def reserve(self, capacity: int) -> None:
        check(self._L.ucfp_corpus_reserve(self._h, capacity))

    @property
    def allocated(self) -> int:
        return int(self._L.ucfp_corpus_capacity(self._h))

    def set_id_base(self, base: int) -> None:
        check(self._L.ucfp_corpus_set_id_base(self._h, base))

    def clear(self) -> None:
        check(self._L.ucfp_corpus_clear(self._h))

    def refresh(self) -> None:
        check(self._L.ucfp_corpus_refresh(self._h))

    def device_rows_ptr(self) -> int:
        return int(self._L.ucfp_corpus_device_rows(self._h) or 0)

    def _outs(self, queries, nq, k, key_dtype_np, key_dtype_t, ids_out, keys_out):
        if ids_out is None:
            if _is_device(queries):
                ids_out = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
                keys_out = torch.empty((nq, k), dtype=key_dtype_t, device=queries.device)
            else:
                ids_out = np.empty((nq, k), dtype=np.uint64)
                keys_out = np.empty((nq, k), dtype=key_dtype_np)
        return ids_out, keys_out

    def scan_hamming(self, queries, k: int, ids_out=None, dist_out=None):
        nq = queries.shape[0]
        ids_out, dist_out = self._outs(queries, nq, k, np.uint32, torch.int32 if torch else None, ids_out, dist_out)
        check(self._L.ucfp_scan_hamming(self._h, _ptr(queries), nq, k, _ptr(ids_out), _ptr(dist_out)))
        return ids_out, dist_out

    def scan_jaccard(self, queries, k: int, ids_out=None, matches_out=None):
        nq = queries.shape[0]
        ids_out, matches_out = self._outs(queries, nq, k, np.uint32, torch.int32 if torch else None, ids_out, matches_out)
        check(self._L.ucfp_scan_jaccard(self._h, _ptr(queries), nq, k, _ptr(ids_out), _ptr(matches_out)))
        return ids_out, matches_out

    def scan_cosine(self, queries, k: int, ids_out=None, score_out=None):
        nq = queries.shape[0]
        ids_out, score_out = self._outs(queries, nq, k, np.float32, torch.float32 if torch else None, ids_out, score_out)
        check(self._L.ucfp_scan_cosine(self._h, _ptr(queries), nq, k, _ptr(ids_out), _ptr(score_out)))
        return ids_out, score_out

    def scan_multihash(self, queries, k_prime: int, k: int, cfg=None, ids_out=None, score_out=None):
        """Re-rank (spec section 10): Hamming top-k' on the PHash global hash, blended global + block score, best k.
        queries: (nq, 51) u64 bundles.  MULTIHASH corpora only."""
        nq = queries.shape[0]
        ids_out, score_out = self._outs(queries, nq, k, np.float32, torch.float32 if torch else None, ids_out, score_out)
        c = _ffi.MultiHashConfig.of(cfg)
        check(self._L.ucfp_scan_multihash(self._h, _ptr(queries), nq, k_prime, k, C.byref(c), _ptr(ids_out), _ptr(score_out)))
        return ids_out, score_out

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.ucfp_corpus_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_KEY_DTYPE = {_ffi.KIND_HAMMING64: np.uint32, _ffi.KIND_MINHASH128: np.uint32, _ffi.KIND_COSINE: np.float32}


class Batcher:
    """ucfp_batcher: coalesces single-query calls from many host threads into batched scans.  `query()` blocks and is
    thread-safe (ctypes releases the GIL for the duration of the call)."""

    def __init__(self, corpus: Corpus, max_batch: int = 512, max_delay_us: int = 200):
        self.corpus, self._L = corpus, corpus._L
        h = C.c_void_p()
        check(self._L.ucfp_batcher_create(corpus._h, max_batch, max_delay_us, C.byref(h)))
        self._h = h

    def query(self, query, k: int):
        q = np.ascontiguousarray(query)
        ids = np.empty(k, dtype=np.uint64)
        keys = np.empty(k, dtype=_KEY_DTYPE[self.corpus.kind])
        check(self._L.ucfp_batcher_query(self._h, _ptr(q), k, _ptr(ids), _ptr(keys)))
        return ids, keys

    def stats(self):
        """-> (queries served, batches scanned, largest batch)"""
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        check(self._L.ucfp_batcher_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.ucfp_batcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """ucfp_group: record-range shards over the GPUs of one box, one NCCL all-gather of packed top-k records per scan.

    Group.local(devices)            one process drives several GPUs (ncclCommInitAll); contexts via .ctx(i)
    Group.join(ctx, id, rank, world) one process per GPU; `id` = Group.unique_id() of rank 0, shipped by the caller"""

    def __init__(self, handle, L):
        self._h, self._L = handle, L
        self._ctxs = {}

    @classmethod
    def local(cls, devices):
        L = _ffi.lib()
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        check(L.ucfp_group_create(arr, len(devices), C.byref(h)))
        g = cls(h, L)
        g.devices = list(devices)
        return g

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_ffi.lib().ucfp_group_unique_id(buf))
        return buf.raw

    @classmethod
    def join(cls, ctx: Context, unique_id: bytes, rank: int, world: int):
        L = _ffi.lib()
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        check(L.ucfp_group_join(ctx._h, buf, rank, world, C.byref(h)))
        g = cls(h, L)
        g.devices = [ctx.device]
        g._ctxs[0] = ctx
        return g

    @property
    def local_size(self) -> int:
        return int(self._L.ucfp_group_local_size(self._h))

    @property
    def world_size(self) -> int:
        return int(self._L.ucfp_group_world_size(self._h))

    def ctx(self, local_rank: int) -> Context:
        """Context of a local rank (owned by the group; pooled mode)."""
        if local_rank not in self._ctxs:
            raw = self._L.ucfp_group_ctx(self._h, local_rank)
            if not raw:
                raise IndexError(local_rank)
            self._ctxs[local_rank] = Context(self.devices[local_rank], use_torch_stream=False, _borrowed=raw)
        return self._ctxs[local_rank]

    def _scan(self, fn, corpora, queries, k, key_np, key_t, ids_out, keys_out):
        nq = queries.shape[0]
        if ids_out is None:
            if _is_device(queries):
                ids_out = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
                keys_out = torch.empty((nq, k), dtype=key_t, device=queries.device)
            else:
                ids_out, keys_out = np.empty((nq, k), dtype=np.uint64), np.empty((nq, k), dtype=key_np)
        arr = (C.c_void_p * len(corpora))(*[c._h for c in corpora])
        check(fn(self._h, arr, _ptr(queries), nq, k, _ptr(ids_out), _ptr(keys_out)))
        return ids_out, keys_out

    def scan_hamming(self, corpora, queries, k, ids_out=None, dist_out=None):
        return self._scan(self._L.ucfp_group_scan_hamming, corpora, queries, k, np.uint32, torch.int32 if torch else None, ids_out, dist_out)

    def scan_jaccard(self, corpora, queries, k, ids_out=None, matches_out=None):
        return self._scan(self._L.ucfp_group_scan_jaccard, corpora, queries, k, np.uint32, torch.int32 if torch else None, ids_out, matches_out)

    def scan_cosine(self, corpora, queries, k, ids_out=None, score_out=None):
        return self._scan(self._L.ucfp_group_scan_cosine, corpora, queries, k, np.float32, torch.float32 if torch else None, ids_out, score_out)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.ucfp_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
