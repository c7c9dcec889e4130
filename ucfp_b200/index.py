"""GPU index backend -- mirror of the reference's `trait IndexBackend` (src/index/mod.rs:17-78) for the scan
path, plus the two entry points the reference lacks (Hamming / Jaccard k-NN).

`GpuIndexBackend.knn` has the semantics of EmbeddedBackend::knn (src/index/embedded/mod.rs:268-360):
per-tenant brute-force cosine, `k == 0` or an empty / zero-norm query -> [], rows whose dimension differs
from the query's are invisible (:307), hits are Vector-sourced and sorted by score descending.  Storage
(redb), BM25 and metadata stay on the host and are out of scope here; `upsert` only mirrors the fields the
scans need into HBM (the corpus is a cache of what redb holds).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from .core import Error, Hit, HitSource, Record
from .image import ALGORITHM_MULTIHASH, global_hash_of, minhash_payload_of
from .runtime import Context, Corpus

_GROW = 2.0


class _Shelf:
    """One (tenant, kind, dim) corpus with amortised growth (HBM rows are append-only; delete rebuilds)."""

    def __init__(self, ctx: Context, kind: int, dim: int, np_dtype, width: int):
        self.ctx, self.kind, self.dim, self.np_dtype, self.width = ctx, kind, dim, np_dtype, width
        self.rows: Dict[int, np.ndarray] = {}     # record_id -> host copy (source of truth for rebuilds)
        self.corpus: Optional[Corpus] = None
        self.dirty = True

    def put(self, rid: int, row: np.ndarray) -> None:
        self.rows[rid] = np.ascontiguousarray(row, dtype=self.np_dtype).reshape(self.width)
        self.dirty = True

    def drop(self, rid: int) -> None:
        if self.rows.pop(rid, None) is not None:
            self.dirty = True

    def resident(self) -> Optional[Corpus]:
        if not self.rows:
            return None
        if self.dirty:
            if self.corpus is not None:
                self.corpus.close()
            ids = np.fromiter(self.rows.keys(), dtype=np.uint64, count=len(self.rows))
            mat = np.stack([self.rows[int(i)] for i in ids]) if self.width > 1 else \
                np.array([self.rows[int(i)][0] for i in ids], dtype=self.np_dtype)
            self.corpus = Corpus(self.ctx, self.kind, max(int(len(ids) * _GROW), 1024), dim=self.dim)
            self.corpus.append(mat, ids)
            self.dirty = False
        return self.corpus


class GpuIndexBackend:
    def __init__(self, ctx: Optional[Context] = None, device: int = 0):
        self.ctx = ctx or Context(device)
        self._vec: Dict[Tuple[int, int], _Shelf] = {}     # (tenant, dim)
        self._ham: Dict[Tuple[int, str], _Shelf] = {}     # (tenant, algorithm)
        self._mh: Dict[int, _Shelf] = {}                  # tenant

    # ---- IndexBackend::upsert / delete (src/index/mod.rs:20-25) ------------------------------------
    def upsert(self, batch: Sequence[Record]) -> None:
        for r in batch:
            self.delete(r.tenant_id, [r.record_id])  # insert-or-replace by (tenant_id, record_id)
            if r.embedding is not None and len(r.embedding) > 0:
                dim = len(r.embedding)
                shelf = self._vec.setdefault((r.tenant_id, dim), _Shelf(self.ctx, _ffi.KIND_COSINE, dim, np.float32, dim))
                shelf.put(r.record_id, np.asarray(r.embedding, dtype=np.float32))
            if r.algorithm.startswith("imgfprint-") and len(r.fingerprint) in (168, 536):
                shelf = self._ham.setdefault((r.tenant_id, r.algorithm), _Shelf(self.ctx, _ffi.KIND_HAMMING64, 0, np.uint64, 1))
                shelf.put(r.record_id, np.array([global_hash_of(r.fingerprint, r.algorithm)], dtype=np.uint64))
            if r.algorithm == "minhash-h128" and len(r.fingerprint) == 1032:
                shelf = self._mh.setdefault(r.tenant_id, _Shelf(self.ctx, _ffi.KIND_MINHASH128, 0, np.uint64, 128))
                shelf.put(r.record_id, minhash_payload_of(r.fingerprint))

    def delete(self, tenant_id: int, ids: Sequence[int]) -> None:
        """Idempotent: missing ids are ignored (src/index/mod.rs:23-25)."""
        for (t, _), shelf in list(self._vec.items()) + list(self._ham.items()):
            if t == tenant_id:
                for i in ids:
                    shelf.drop(i)
        if tenant_id in self._mh:
            for i in ids:
                self._mh[tenant_id].drop(i)

    def hydrate_fingerprints(self, tenant_id: int, algorithm: str, record_ids, blobs: bytes) -> None:
        """Bulk load: `blobs` = equally sized fingerprint blobs back to back, as a range scan of the
        redb fingerprints table yields them (src/index/embedded/mod.rs:37-43).  One strided copy per call instead of
        one upsert per record."""
        ids = np.ascontiguousarray(record_ids, dtype=np.uint64)
        n = len(ids)
        if n == 0:
            return
        size = len(blobs) // n
        buf = np.frombuffer(blobs, dtype=np.uint8)
        if algorithm == "minhash-h128":
            if size != 1032:
                raise Error("Incompatible", f"MinHashSig<128> blobs must be 1032 bytes, got {size}")
            shelf = self._mh.setdefault(tenant_id, _Shelf(self.ctx, _ffi.KIND_MINHASH128, 0, np.uint64, 128))
            off, view = 8, buf.reshape(n, size)[:, 8:].copy().view(np.uint64)
        else:
            off = 232 if algorithm == ALGORITHM_MULTIHASH else 32
            if size != (536 if algorithm == ALGORITHM_MULTIHASH else 168):
                raise Error("Incompatible", f"{algorithm} blobs have the wrong size {size}")
            shelf = self._ham.setdefault((tenant_id, algorithm), _Shelf(self.ctx, _ffi.KIND_HAMMING64, 0, np.uint64, 1))
            view = buf.reshape(n, size)[:, off:off + 8].copy().view(np.uint64)
        for rid, row in zip(ids, view):          # host copy stays the source of truth for deletes/rebuilds
            shelf.rows[int(rid)] = np.ascontiguousarray(row).reshape(shelf.width)
        # the HBM mirror itself is filled by one strided copy straight from the blob run
        if shelf.corpus is not None:
            shelf.corpus.close()
        shelf.corpus = Corpus(self.ctx, shelf.kind, max(int(len(shelf.rows) * _GROW), 1024), dim=shelf.dim)
        if len(shelf.rows) == n:
            shelf.corpus.append_strided(buf, size, off, n, ids)
            shelf.dirty = False
        else:
            shelf.dirty = True

    def flush(self) -> None:
        self.ctx.synchronize()

    # ---- IndexBackend::knn (src/index/mod.rs:29-35) -------------------------------------------------
    def knn(self, tenant_id: int, query: Sequence[float], k: int, _filter: Optional[bytes] = None) -> List[Hit]:
        return self.knn_batch(tenant_id, [query], k)[0] if len(query) and k else []

    def knn_batch(self, tenant_id: int, queries, k: int) -> List[List[Hit]]:
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] == 0 or k == 0:
            return [[] for _ in range(len(q))]
        shelf = self._vec.get((tenant_id, q.shape[1]))
        corpus = shelf.resident() if shelf else None
        if corpus is None:
            return [[] for _ in range(len(q))]
        ids, scores = corpus.scan_cosine(q, k)
        return [[Hit(tenant_id, int(i), float(s), HitSource.VECTOR) for i, s in zip(ri, rs) if i != _ffi.ID_NONE]
                for ri, rs in zip(ids, scores)]

    # ---- new: HashIndex (SURVEY 8b "new API the reference lacks") ----------------------------------
    def hamming_knn(self, tenant_id: int, algorithm: str, code: int, k: int) -> List[Hit]:
        """Hit.score = 1 - dist / 64, source = Vector."""
        shelf = self._ham.get((tenant_id, algorithm))
        corpus = shelf.resident() if shelf else None
        if corpus is None or k == 0:
            return []
        ids, dist = corpus.scan_hamming(np.array([code], dtype=np.uint64), k)
        return [Hit(tenant_id, int(i), 1.0 - float(d) / 64.0, HitSource.VECTOR) for i, d in zip(ids[0], dist[0]) if i != _ffi.ID_NONE]

    def jaccard_knn(self, tenant_id: int, signature, k: int) -> List[Hit]:
        """Hit.score = matches / 128, source = Vector."""
        shelf = self._mh.get(tenant_id)
        corpus = shelf.resident() if shelf else None
        if corpus is None or k == 0:
            return []
        sig = np.ascontiguousarray(signature, dtype=np.uint64).reshape(1, 128)
        ids, m = corpus.scan_jaccard(sig, k)
        return [Hit(tenant_id, int(i), float(x) / 128.0, HitSource.VECTOR) for i, x in zip(ids[0], m[0]) if i != _ffi.ID_NONE]

    def bm25(self, tenant_id: int, terms, k: int, _filter=None) -> List[Hit]:
        raise Error("Unsupported", "bm25 stays on the host backend (out of scope for the GPU hot path)")
