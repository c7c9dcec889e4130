"""GPU index backend -- mirror of the reference's `trait IndexBackend` (src/index/mod.rs:17-78) for the scan
path, plus the entry points the reference lacks (Hamming / Jaccard k-NN and the block-hash re-rank).

`GpuIndexBackend.knn` has the semantics of EmbeddedBackend::knn (src/index/embedded/mod.rs:268-360):
per-tenant brute-force cosine, `k == 0` or an empty / zero-norm query -> [], rows whose dimension differs
from the query's are invisible (:307), hits are Vector-sourced and sorted by score descending, at most min(k, N) hits.
Storage (redb), BM25 and metadata stay on the host and are out of scope here: the corpora are the HBM mirror of what
redb holds.  `upsert` / `delete` go straight to ucfp_corpus_upsert / ucfp_corpus_delete (insert-or-replace and
idempotent delete by record id on the device): nothing is rebuilt, no host copy of the rows is kept.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from .core import Error, Hit, HitSource, Record
from .image import MULTIHASH_TAGS, SINGLE_HASH_TAGS, bundle_words_of, global_hash_of, minhash_payload_of
from .runtime import Context, Corpus

K_LIMIT = {_ffi.KIND_COSINE: 1024, _ffi.KIND_HAMMING64: 2048, _ffi.KIND_MINHASH128: 2048, _ffi.KIND_MULTIHASH: 2048}   # include/ucfp_cuda.h
_FIRST_CAPACITY = 1024


class _Shelf:
    """One (tenant, kind, dim) corpus.  It grows by itself (ucfp_corpus_upsert reallocates a full corpus)."""

    def __init__(self, ctx: Context, kind: int, dim: int, np_dtype, width: int):
        self.ctx, self.kind, self.dim, self.np_dtype, self.width = ctx, kind, dim, np_dtype, width
        self.corpus: Optional[Corpus] = None

    def _ensure(self, rows_hint: int) -> Corpus:
        if self.corpus is None:
            self.corpus = Corpus(self.ctx, self.kind, max(_FIRST_CAPACITY, rows_hint), dim=self.dim)
        return self.corpus

    def upsert(self, ids: np.ndarray, rows: np.ndarray) -> None:
        self._ensure(len(ids)).upsert(np.ascontiguousarray(ids, dtype=np.uint64),
                                      np.ascontiguousarray(rows, dtype=self.np_dtype).reshape(len(ids), self.width))

    def delete(self, ids: np.ndarray) -> None:
        if self.corpus is not None and len(self.corpus):
            self.corpus.delete(ids)

    def resident(self) -> Optional[Corpus]:
        return self.corpus if self.corpus is not None and len(self.corpus) else None

    def clamp_k(self, k: int) -> int:
        """EmbeddedBackend::knn returns min(k, N) hits; the scan kernels cap k per call (include/ucfp_cuda.h)."""
        k = min(int(k), len(self.corpus))
        if k > K_LIMIT[self.kind]:
            raise Error("Unsupported", f"k = {k} exceeds the scan limit of {K_LIMIT[self.kind]} results per query")
        return k


class GpuIndexBackend:
    def __init__(self, ctx: Optional[Context] = None, device: int = 0):
        self.ctx = ctx or Context(device)
        self._vec: Dict[Tuple[int, int], _Shelf] = {}     # (tenant, dim)
        self._ham: Dict[Tuple[int, str], _Shelf] = {}     # (tenant, algorithm): one u64 code per record
        self._multi: Dict[Tuple[int, str], _Shelf] = {}   # (tenant, algorithm): 51-word bundles, for the block-hash re-rank
        self._mh: Dict[int, _Shelf] = {}                  # tenant

    def _shelves_of(self, tenant_id: int):
        for table in (self._vec, self._ham, self._multi):
            for (t, _), shelf in table.items():
                if t == tenant_id:
                    yield shelf
        if tenant_id in self._mh:
            yield self._mh[tenant_id]

    # ---- IndexBackend::upsert / delete (src/index/mod.rs:20-25) ------------------------------------
    def upsert(self, batch: Sequence[Record]) -> None:
        """Insert-or-replace by (tenant_id, record_id).  A record that changes shape (e.g. gets a new embedding dimension
        or algorithm) leaves its old shelves first, as the redb row it replaces would."""
        by_tenant = defaultdict(list)
        for r in batch:
            by_tenant[r.tenant_id].append(r.record_id)
        groups = defaultdict(lambda: ([], []))             # shelf -> (ids, rows)
        for r in batch:
            if r.embedding is not None and len(r.embedding) > 0:
                dim = len(r.embedding)
                shelf = self._vec.setdefault((r.tenant_id, dim), _Shelf(self.ctx, _ffi.KIND_COSINE, dim, np.float32, dim))
                groups[shelf][0].append(r.record_id); groups[shelf][1].append(np.asarray(r.embedding, dtype=np.float32))
            if r.algorithm in MULTIHASH_TAGS and len(r.fingerprint) == 536 or r.algorithm in SINGLE_HASH_TAGS and len(r.fingerprint) == 168:
                shelf = self._ham.setdefault((r.tenant_id, r.algorithm), _Shelf(self.ctx, _ffi.KIND_HAMMING64, 0, np.uint64, 1))
                groups[shelf][0].append(r.record_id); groups[shelf][1].append(np.array([global_hash_of(r.fingerprint, r.algorithm)], dtype=np.uint64))
            if r.algorithm in MULTIHASH_TAGS and len(r.fingerprint) == 536:
                shelf = self._multi.setdefault((r.tenant_id, r.algorithm), _Shelf(self.ctx, _ffi.KIND_MULTIHASH, 0, np.uint64, 51))
                groups[shelf][0].append(r.record_id); groups[shelf][1].append(bundle_words_of(r.fingerprint))
            if r.algorithm == "minhash-h128" and len(r.fingerprint) == 1032:
                shelf = self._mh.setdefault(r.tenant_id, _Shelf(self.ctx, _ffi.KIND_MINHASH128, 0, np.uint64, 128))
                groups[shelf][0].append(r.record_id); groups[shelf][1].append(minhash_payload_of(r.fingerprint))
        for tenant_id, ids in by_tenant.items():            # replaced records leave the shelves they no longer belong to
            ids = np.asarray(ids, dtype=np.uint64)
            for shelf in self._shelves_of(tenant_id):
                if shelf not in groups:
                    shelf.delete(ids)
                else:
                    stay = np.asarray(groups[shelf][0], dtype=np.uint64)
                    gone = np.setdiff1d(ids, stay)
                    if len(gone):
                        shelf.delete(gone)
        for shelf, (ids, rows) in groups.items():
            shelf.upsert(np.asarray(ids, dtype=np.uint64), np.stack(rows))

    def delete(self, tenant_id: int, ids: Sequence[int]) -> None:
        """Idempotent: missing ids are ignored (src/index/mod.rs:23-25)."""
        ids = np.asarray(list(ids), dtype=np.uint64)
        if len(ids):
            for shelf in self._shelves_of(tenant_id):
                shelf.delete(ids)

    def hydrate_fingerprints(self, tenant_id: int, algorithm: str, record_ids, blobs: bytes) -> None:
        """Bulk load: `blobs` = equally sized fingerprint blobs back to back, as a range scan of the
        redb fingerprints table yields them (src/index/embedded/mod.rs:37-43).  One strided copy per corpus, no per-record
        work; into a shelf that already holds rows the fields are extracted on the host and upserted."""
        ids = np.ascontiguousarray(record_ids, dtype=np.uint64)
        n = len(ids)
        if n == 0:
            return
        size = len(blobs) // n
        buf = np.frombuffer(blobs, dtype=np.uint8)
        plans = []                                           # (shelf, field offset, host view for the upsert path)
        if algorithm == "minhash-h128":
            if size != 1032:
                raise Error("Incompatible", f"MinHashSig<128> blobs must be 1032 bytes, got {size}")
            shelf = self._mh.setdefault(tenant_id, _Shelf(self.ctx, _ffi.KIND_MINHASH128, 0, np.uint64, 128))
            plans.append((shelf, 8, lambda: buf.reshape(n, size)[:, 8:].copy().view(np.uint64)))
        elif algorithm in MULTIHASH_TAGS or algorithm in SINGLE_HASH_TAGS:
            multi = algorithm in MULTIHASH_TAGS
            off = 232 if multi else 32
            if size != (536 if multi else 168):
                raise Error("Incompatible", f"{algorithm} blobs have the wrong size {size}")
            shelf = self._ham.setdefault((tenant_id, algorithm), _Shelf(self.ctx, _ffi.KIND_HAMMING64, 0, np.uint64, 1))
            plans.append((shelf, off, lambda: buf.reshape(n, size)[:, off:off + 8].copy().view(np.uint64)))
            if multi:
                shelf = self._multi.setdefault((tenant_id, algorithm), _Shelf(self.ctx, _ffi.KIND_MULTIHASH, 0, np.uint64, 51))
                plans.append((shelf, 0, lambda: np.concatenate([buf.reshape(n, size)[:, 64 + 168 * a: 200 + 168 * a] for a in range(3)], axis=1).copy().view(np.uint64)))
        else:
            raise Error("Unsupported", f"no scan corpus for algorithm {algorithm!r}")
        for shelf, off, host_rows in plans:
            if shelf.resident() is None:
                corpus = shelf._ensure(n)
                corpus.reserve(n)
                corpus.append_strided(buf, size, off, n, ids)
            else:
                shelf.upsert(ids, host_rows())

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
        ids, scores = corpus.scan_cosine(q, shelf.clamp_k(k))
        return [[Hit(tenant_id, int(i), float(s), HitSource.VECTOR) for i, s in zip(ri, rs) if i != _ffi.ID_NONE]
                for ri, rs in zip(ids, scores)]

    # ---- new: HashIndex (SURVEY 8b "new API the reference lacks") ----------------------------------
    def hamming_knn(self, tenant_id: int, algorithm: str, code: int, k: int) -> List[Hit]:
        """Hit.score = 1 - dist / 64, source = Vector."""
        shelf = self._ham.get((tenant_id, algorithm))
        corpus = shelf.resident() if shelf else None
        if corpus is None or k == 0:
            return []
        ids, dist = corpus.scan_hamming(np.array([code], dtype=np.uint64), shelf.clamp_k(k))
        return [Hit(tenant_id, int(i), 1.0 - float(d) / 64.0, HitSource.VECTOR) for i, d in zip(ids[0], dist[0]) if i != _ffi.ID_NONE]

    def jaccard_knn(self, tenant_id: int, signature, k: int) -> List[Hit]:
        """Hit.score = matches / 128, source = Vector."""
        shelf = self._mh.get(tenant_id)
        corpus = shelf.resident() if shelf else None
        if corpus is None or k == 0:
            return []
        sig = np.ascontiguousarray(signature, dtype=np.uint64).reshape(1, 128)
        ids, m = corpus.scan_jaccard(sig, shelf.clamp_k(k))
        return [Hit(tenant_id, int(i), float(x) / 128.0, HitSource.VECTOR) for i, x in zip(ids[0], m[0]) if i != _ffi.ID_NONE]

    def multihash_knn(self, tenant_id: int, algorithm: str, bundle: bytes, k: int, k_prime: Optional[int] = None, config=None) -> List[Hit]:
        """Block-hash-aware search (docs/HASH_SPEC.md section 10): the k' nearest PHash global hashes re-ranked by
        the blended global + block similarity of the whole 536-byte bundle.  Hit.score = that similarity; `config` = the
        reference's MultiHashConfigDto fields (src/server/dto.rs:462-480), kebab- or snake-case."""
        shelf = self._multi.get((tenant_id, algorithm))
        corpus = shelf.resident() if shelf else None
        if corpus is None or k == 0:
            return []
        k = shelf.clamp_k(k)
        kp = min(max(k_prime or 8 * k, k), len(corpus), 2048)
        cfg = {key.replace("-", "_"): v for key, v in (config or {}).items() if v is not None}
        ids, sc = corpus.scan_multihash(bundle_words_of(bundle).reshape(1, 51), kp, k, cfg)
        return [Hit(tenant_id, int(i), float(s), HitSource.VECTOR) for i, s in zip(ids[0], sc[0]) if i != _ffi.ID_NONE]

    def bm25(self, tenant_id: int, terms, k: int, _filter=None) -> List[Hit]:
        raise Error("Unsupported", "bm25 stays on the host backend (out of scope for the GPU hot path)")
