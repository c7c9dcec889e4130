"""`POST /v1/query` mirror (ucfp_b200/server.py): request parsing and response shape follow the reference
(src/server/dto.rs:75-116, src/server/handlers.rs:139-197); the `hash` / `signature` kinds are this project's extension
(SURVEY 8f N2).  A recording fake stands in for the index, so this runs without a GPU."""
import json

import pytest

from ucfp_b200 import image

from ucfp_b200 import Error, Hit, HitSource, Modality, server


class FakeIndex:
    def __init__(self):
        self.calls = []

    def knn(self, tenant_id, vector, k, filt=None):
        self.calls.append(("knn", tenant_id, list(vector), k))
        return [Hit(tenant_id, 300, 0.98, HitSource.VECTOR), Hit(tenant_id, 100, 0.70, HitSource.VECTOR),
                Hit(tenant_id, 200, 0.69, HitSource.VECTOR)][:k]

    def hamming_knn(self, tenant_id, algorithm, code, k):
        self.calls.append(("hamming", tenant_id, algorithm, code, k))
        return [Hit(tenant_id, 7, 1.0, HitSource.VECTOR), Hit(tenant_id, 9, 1.0 - 3 / 64.0, HitSource.VECTOR)][:k]

    def jaccard_knn(self, tenant_id, signature, k):
        self.calls.append(("jaccard", tenant_id, list(signature), k))
        return [Hit(tenant_id, 5, 120 / 128.0, HitSource.VECTOR)][:k]


def test_reference_vector_request_and_response_shape():
    idx = FakeIndex()
    out = server.query(idx, json.dumps({"tenant_id": 7, "modality": "Text", "k": 2, "vector": [0.6, 0.6, 0]}))
    assert idx.calls == [("knn", 7, [0.6, 0.6, 0.0], 2)]
    # HitOut: Option fields that are None and the empty term_hits are omitted (serde skip_serializing_if)
    assert out == {"hits": [{"tenant_id": 7, "record_id": 300, "score": 0.98, "source": "vector"},
                            {"tenant_id": 7, "record_id": 100, "score": 0.70, "source": "vector"}]}


def test_k_defaults_to_10_and_is_at_least_1():
    q = server.parse_query_request({"tenant_id": 1, "modality": "Image", "vector": [1.0]})
    assert q.k == 10 and q.rrf_k == 60 and q.terms == [] and q.filter is None and q.modality is Modality.IMAGE
    assert server.parse_query_request({"tenant_id": 1, "modality": "Image", "k": 0, "vector": [1.0]}).k == 1


def test_explain_parameter_spellings():
    assert [server.parse_explain(v) for v in ("1", "true", "yes", "0", "no", "TRUE", None)] == [True, True, True, False, False, False, False]
    assert server.parse_query_request({"tenant_id": 1, "modality": "Text", "vector": [1]}, explain="yes").explain is True


def test_hash_query_goes_to_the_hamming_arm():
    idx = FakeIndex()
    out = server.query(idx, {"tenant_id": 3, "modality": "Image", "k": 5, "hash": "0xFEEDFACECAFEBEEF", "algorithm": "imgfprint-phash-v1"})
    assert idx.calls == [("hamming", 3, "imgfprint-phash-v1", 0xFEEDFACECAFEBEEF, 5)]
    assert [h["record_id"] for h in out["hits"]] == [7, 9] and out["hits"][1]["score"] == 1.0 - 3 / 64.0
    idx = FakeIndex()
    server.query(idx, {"tenant_id": 3, "modality": "Image", "hash": 2**64 - 1})          # integer form, default algorithm = multi bundle
    assert idx.calls == [("hamming", 3, image.ALGORITHM_MULTIHASH, 2**64 - 1, 10)]


def test_signature_query_goes_to_the_jaccard_arm():
    idx = FakeIndex()
    sig = list(range(1, 129))
    out = server.query(idx, {"tenant_id": 4, "modality": "Text", "k": 1, "signature": sig})
    assert idx.calls == [("jaccard", 4, sig, 1)] and out["hits"][0]["score"] == 120 / 128.0


@pytest.mark.parametrize("body", [
    b"not json", "[]", {"modality": "Text", "vector": [1]}, {"tenant_id": 1, "vector": [1]},
    {"tenant_id": 1, "modality": "text", "vector": [1]},                 # serde variant names are capitalised
    {"tenant_id": 2**32, "modality": "Text", "vector": [1]}, {"tenant_id": -1, "modality": "Text", "vector": [1]},
    {"tenant_id": 1, "modality": "Text"},                                # the reference requires `vector`
    {"tenant_id": 1, "modality": "Text", "vector": [1], "hash": 5},      # one query kind only
    {"tenant_id": 1, "modality": "Text", "vector": "abc"}, {"tenant_id": 1, "modality": "Text", "hash": 2**64},
    {"tenant_id": 1, "modality": "Text", "hash": "zz"}, {"tenant_id": 1, "modality": "Text", "signature": [1, 2, 3]},
    {"tenant_id": 1, "modality": "Text", "hash": 1, "algorithm": ""}, {"tenant_id": True, "modality": "Text", "vector": [1]},
])
def test_malformed_requests_are_rejected(body):
    with pytest.raises(Error) as e:
        server.parse_query_request(body)
    assert e.value.kind == "BadRequest"
