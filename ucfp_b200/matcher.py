"""Matcher -- mirror of src/matcher/mod.rs: `search` dispatch (:140-207) and RRF fusion (:22-98).

Only the vector arm reaches the GPU (`IndexBackend::knn`); fusion and truncation are host-side and tiny
(<= 2k items), kept here so that callers of `Matcher::search` find the same behaviour."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .core import Hit, HitSource, Query


def rrf_with_sources(rankings: Sequence[Sequence[Hit]], sources: Sequence[HitSource], rrf_k: int) -> List[Hit]:
    """src/matcher/mod.rs:32-98, f32 arithmetic included."""
    denom = np.float32(rrf_k)
    acc: Dict[Tuple[int, int], list] = {}
    for i, ranking in enumerate(rankings):
        src = sources[i] if i < len(sources) else (ranking[0].source if ranking else HitSource.FUSED)
        for rank0, hit in enumerate(ranking):
            rank1 = rank0 + 1
            inc = np.float32(1.0) / (denom + np.float32(rank1))
            e = acc.setdefault((hit.tenant_id, hit.record_id), [None, None, None, None])
            if src == HitSource.VECTOR:
                e[0] = np.float32((e[0] or np.float32(0)) + inc)
                e[2] = e[2] if e[2] is not None else rank1
            elif src == HitSource.BM25:
                e[1] = np.float32((e[1] or np.float32(0)) + inc)
                e[3] = e[3] if e[3] is not None else rank1
            else:
                e[0] = np.float32((e[0] or np.float32(0)) + inc)
    out = [Hit(t, r, float(np.float32((vs or np.float32(0)) + (bs or np.float32(0)))), HitSource.FUSED,
               None if vs is None else float(vs), None if bs is None else float(bs), vr, br)
           for (t, r), (vs, bs, vr, br) in acc.items()]
    out.sort(key=lambda h: -h.score)
    return out


def rrf(rankings: Sequence[Sequence[Hit]], rrf_k: int) -> List[Hit]:
    """src/matcher/mod.rs:22-24."""
    return rrf_with_sources(rankings, [], rrf_k)


class Matcher:
    def __init__(self, index, reranker=None):
        self.index, self.reranker = index, reranker

    def search(self, q: Query) -> List[Hit]:
        """src/matcher/mod.rs:140-207."""
        has_terms = len(q.terms) > 0
        if q.vector is not None and has_terms:
            vec_hits = self.index.knn(q.tenant_id, q.vector, q.k, q.filter)
            bm_hits = self.index.bm25(q.tenant_id, q.terms, q.k, q.filter)
            fused = rrf_with_sources([vec_hits, bm_hits], [HitSource.VECTOR, HitSource.BM25], q.rrf_k)
        elif q.vector is not None:
            fused = self.index.knn(q.tenant_id, q.vector, q.k, q.filter)
        elif has_terms:
            fused = self.index.bm25(q.tenant_id, q.terms, q.k, q.filter)
        elif q.hash is not None:           # new arms (absent upstream): the HashIndex queries of SURVEY 8b
            fused = self.index.hamming_knn(q.tenant_id, q.hash_algorithm, q.hash, q.k)
        elif q.signature is not None:
            fused = self.index.jaccard_knn(q.tenant_id, q.signature, q.k)
        else:
            fused = []
        fused = fused[: q.k]
        if self.reranker is not None:
            fused = self.reranker.rerank(q, fused)
        return fused
