"""`POST /v1/query` -- host-side mirror of the reference's request/response shapes and handler
(src/server/dto.rs:75-116 `QueryRequest` / `QueryResponse` / `HitOut`, src/server/handlers.rs:139-197 `query`,
`parse_explain`, `hit_source_str`), extended with the two query kinds the reference stores fingerprints for but cannot
ask: a 64-bit perceptual-hash code and a MinHash-128 signature.

Reference body:            {"tenant_id": u32, "modality": "Audio"|"Image"|"Text", "k": usize = 10, "vector": [f32]}
Extension (exactly one of `vector`, `hash`, `signature`):
    "hash": u64 as a JSON integer or a "0x..." / decimal string, with "algorithm": the stored fingerprint's algorithm tag
            (src/modality/image.rs:38-46, default the multi-hash bundle, whose PHash global code is searched)
    "signature": 128 u64 slots of a txtfp MinHashSig<128> (src/modality/text.rs:200-204)
Scores: cosine as the reference; Hamming 1 - dist / 64; Jaccard matches / 128.  No HTTP server lives here (out of scope):
`query()` is what a route handler would call with the decoded body and the `?explain=` parameter."""
from __future__ import annotations

import json
from typing import Any, Dict, Optional, Union

from .core import Error, Hit, Modality, Query
from .matcher import Matcher

DEFAULT_K = 10                                    # dto.rs:86 default_k
_MODALITY = {"Audio": Modality.AUDIO, "Image": Modality.IMAGE, "Text": Modality.TEXT}   # serde: variant names, core/mod.rs:17-25
from .image import ALGORITHM_MULTIHASH as DEFAULT_HASH_ALGORITHM  # the multi bundle (image.rs:38), under the tag this build stamps


def parse_explain(value: Optional[str]) -> bool:
    """handlers.rs:139-141: `?explain=1|true|yes`."""
    return value in ("1", "true", "yes")


def _uint(v: Any, bits: int, what: str) -> int:
    if isinstance(v, bool) or not isinstance(v, (int, str)):
        raise Error("BadRequest", f"{what} must be an unsigned {bits}-bit integer")
    try:
        x = int(v, 0) if isinstance(v, str) else v
    except ValueError:
        raise Error("BadRequest", f"{what} must be an unsigned {bits}-bit integer") from None
    if not 0 <= x < (1 << bits):
        raise Error("BadRequest", f"{what} out of range for u{bits}")
    return x


def parse_query_request(body: Union[bytes, str, Dict[str, Any]], explain: Optional[str] = None) -> Query:
    """Decoded `QueryRequest` -> `Query`, as handlers.rs:148-159 builds it (k.max(1), rrf_k 60, no terms, no filter)."""
    if isinstance(body, (bytes, str)):
        try:
            body = json.loads(body)
        except ValueError as e:
            raise Error("BadRequest", f"body is not JSON: {e}") from None
    if not isinstance(body, dict):
        raise Error("BadRequest", "body must be a JSON object")
    for key in ("tenant_id", "modality"):
        if key not in body:
            raise Error("BadRequest", f"missing field `{key}`")
    tenant_id = _uint(body["tenant_id"], 32, "tenant_id")
    if body["modality"] not in _MODALITY:
        raise Error("BadRequest", f"unknown variant `{body['modality']}`, expected one of `Audio`, `Image`, `Text`")
    k = _uint(body.get("k", DEFAULT_K), 64, "k")
    kinds = [key for key in ("vector", "hash", "signature") if body.get(key) is not None]
    if len(kinds) != 1:
        raise Error("BadRequest", "exactly one of `vector`, `hash`, `signature` is required" if kinds else "missing field `vector`")
    q = Query(tenant_id=tenant_id, modality=_MODALITY[body["modality"]], k=max(k, 1), rrf_k=60, explain=parse_explain(explain))
    if kinds[0] == "vector":
        v = body["vector"]
        if not isinstance(v, list) or not all(isinstance(x, (int, float)) and not isinstance(x, bool) for x in v):
            raise Error("BadRequest", "vector must be an array of numbers")
        q.vector = [float(x) for x in v]
    elif kinds[0] == "hash":
        q.hash = _uint(body["hash"], 64, "hash")
        algo = body.get("algorithm", DEFAULT_HASH_ALGORITHM)
        if not isinstance(algo, str) or not algo:
            raise Error("BadRequest", "algorithm must be a non-empty string")
        q.hash_algorithm = algo
    else:
        s = body["signature"]
        if not isinstance(s, list) or len(s) != 128:
            raise Error("BadRequest", "signature must be an array of 128 u64 slots")
        q.signature = [_uint(x, 64, "signature slot") for x in s]
    return q


def hit_out(h: Hit) -> Dict[str, Any]:
    """`HitOut` with serde's skip rules (dto.rs:95-116): None options and an empty term_hits are omitted."""
    out: Dict[str, Any] = {"tenant_id": h.tenant_id, "record_id": h.record_id, "score": h.score, "source": h.source.value}
    for key in ("vector_score", "bm25_score", "vector_rank", "bm25_rank"):
        v = getattr(h, key)
        if v is not None:
            out[key] = v
    if h.term_hits:
        out["term_hits"] = [{"term": t.term, "idf": t.idf, "tf": t.tf, "contribution": t.contribution} for t in h.term_hits]
    return out


def query(index, body: Union[bytes, str, Dict[str, Any]], explain: Optional[str] = None) -> Dict[str, Any]:
    """handlers.rs:143-187 without the HTTP and auth layers: parse, `Matcher::search`, map to `QueryResponse`."""
    q = parse_query_request(body, explain)
    return {"hits": [hit_out(h) for h in Matcher(index).search(q)]}
