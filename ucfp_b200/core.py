"""Core data shapes -- mirror of the reference's src/core/mod.rs (Modality :19, Record :34-72,
Hit :108-131, HitSource :135-146, Query :153-189), kept field-for-field so the parity tests read like
the reference's own."""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import List, Optional


class Modality(enum.Enum):
    AUDIO = "audio"
    IMAGE = "image"
    TEXT = "text"


class HitSource(enum.Enum):
    VECTOR = "vector"
    BM25 = "bm25"
    FILTER = "filter"
    RERANKER = "reranker"
    FUSED = "fused"


@dataclass
class Record:
    """src/core/mod.rs:34-72."""
    tenant_id: int
    record_id: int
    modality: Modality
    format_version: int
    algorithm: str
    config_hash: int
    fingerprint: bytes
    embedding: Optional[List[float]] = None
    model_id: Optional[str] = None
    metadata: bytes = b""
    text: Optional[str] = None


@dataclass
class Hit:
    """src/core/mod.rs:108-131."""
    tenant_id: int
    record_id: int
    score: float
    source: HitSource
    vector_score: Optional[float] = None
    bm25_score: Optional[float] = None
    vector_rank: Optional[int] = None
    bm25_rank: Optional[int] = None
    term_hits: list = field(default_factory=list)


@dataclass
class Query:
    """src/core/mod.rs:153-189 (defaults :176-189)."""
    tenant_id: int = 0
    modality: Modality = Modality.TEXT
    k: int = 10
    vector: Optional[List[float]] = None
    terms: List[str] = field(default_factory=list)
    filter: Optional[bytes] = None
    rrf_k: int = 60
    explain: bool = False
    # not in the reference: the two query kinds its stored fingerprints call for
    hash: Optional[int] = None              # 64-bit perceptual-hash code -> Hamming top-k
    hash_algorithm: Optional[str] = None    # algorithm tag of the stored fingerprints to search (image.rs:38-46)
    signature: Optional[List[int]] = None   # MinHash-128 slots -> Jaccard top-k


class Error(Exception):
    """src/error.rs:9-61; `kind` is the variant name."""

    def __init__(self, kind: str, message: str):
        super().__init__(f"{kind}: {message}")
        self.kind = kind
        self.message = message
