"""The host-side mirror of the reference's interfaces, running on the GPU: these read like the reference's own
tests (src/index/embedded/mod.rs:523-589, src/server/tests.rs:239-280,456-532,1167-1208)."""
import io

import numpy as np
import pytest

import oracle
import ucfp_b200
from ucfp_b200 import Error, GpuIndexBackend, HitSource, Matcher, Modality, Query, Record, image

pytestmark = pytest.mark.gpu


def rec(tenant, rid, embedding):
    return Record(tenant_id=tenant, record_id=rid, modality=Modality.IMAGE, format_version=1, algorithm="test", config_hash=0,
                  fingerprint=b"fp", embedding=embedding, model_id="test-model")


def synthetic_png(w, h):
    """src/server/tests.rs:227-235."""
    from PIL import Image
    y, x = np.mgrid[0:h, 0:w]
    arr = np.stack([x % 256, y % 256, np.full_like(x, 128)], -1).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="PNG")
    return buf.getvalue(), arr


def test_upsert_and_knn_round_trip(ctx):
    db = GpuIndexBackend(ctx)
    db.upsert([rec(1, 100, [1.0, 0.0, 0.0]), rec(1, 200, [0.0, 1.0, 0.0]), rec(1, 300, [0.7, 0.7, 0.0])])
    hits = db.knn(1, [0.6, 0.6, 0.0], 2)
    assert len(hits) == 2
    assert hits[0].record_id == 300, "closest match should be 300"
    assert hits[0].score > hits[1].score
    for h in hits:
        assert h.tenant_id == 1 and h.source == HitSource.VECTOR


def test_knn_ignores_other_tenants(ctx):
    db = GpuIndexBackend(ctx)
    db.upsert([rec(1, 1, [1.0, 0.0]), rec(2, 1, [1.0, 0.0])])
    hits = db.knn(1, [1.0, 0.0], 10)
    assert len(hits) == 1 and hits[0].tenant_id == 1


def test_delete_removes_records(ctx):
    db = GpuIndexBackend(ctx)
    db.upsert([rec(1, 1, [1.0, 0.0]), rec(1, 2, [0.0, 1.0])])
    db.delete(1, [1])
    hits = db.knn(1, [1.0, 0.0], 10)
    assert len(hits) == 1 and hits[0].record_id == 2


def test_knn_skips_records_with_no_embedding_or_other_dimension(ctx):
    db = GpuIndexBackend(ctx)
    without = rec(1, 9, None)
    db.upsert([without, rec(1, 10, [1.0, 0.0]), rec(1, 11, [1.0, 0.0, 0.0])])
    hits = db.knn(1, [1.0, 0.0], 10)
    assert len(hits) == 1 and hits[0].record_id == 10
    assert db.knn(1, [], 10) == [] and db.knn(1, [1.0, 0.0], 0) == [] and db.knn(1, [0.0, 0.0], 10) == []


def test_matcher_vector_arm_and_truncate(ctx):
    db = GpuIndexBackend(ctx)
    db.upsert([rec(7, i, [1.0, float(i)]) for i in range(1, 30)])
    hits = Matcher(db).search(Query(tenant_id=7, modality=Modality.IMAGE, k=5, vector=[1.0, 0.0]))
    assert [h.record_id for h in hits] == [1, 2, 3, 4, 5]
    assert Matcher(db).search(Query(tenant_id=7, k=5)) == []


def test_ingest_image_round_trip_multi_and_single(ctx):
    png, arr = synthetic_png(64, 64)
    r = image.fingerprint(png, 9, 42)
    assert r.algorithm == image.ALGORITHM_MULTIHASH == "ucfp-b200-multihash-v1" and r.modality == Modality.IMAGE and r.config_hash == 0
    assert len(r.fingerprint) == 536 and len(r.fingerprint.hex()) == 1072
    words = oracle.image_multihash(arr)
    assert r.fingerprint == image.pack_multihash(image.exact_hash(png), words)
    pre = image.PreprocessConfig()
    for fn, tag, off in ((image.fingerprint_phash, image.ALGORITHM_PHASH, 17), (image.fingerprint_dhash, image.ALGORITHM_DHASH, 34),
                         (image.fingerprint_ahash, image.ALGORITHM_AHASH, 0)):
        s = fn(png, pre, 9, 43)
        assert s.algorithm == tag and len(s.fingerprint) == 168
        assert s.fingerprint == image.pack_image_fingerprint(image.exact_hash(png), words[off:off + 17])


def test_garbage_bytes_and_guards_are_modality_errors(ctx):
    with pytest.raises(Error) as e:
        image.fingerprint(b"not an image", 1, 1)
    assert e.value.kind == "Modality"
    png, _ = synthetic_png(16, 16)
    with pytest.raises(Error):                                    # below min_dimension 32
        image.fingerprint(png, 1, 1)
    big, _ = synthetic_png(64, 64)
    with pytest.raises(Error):
        image.fingerprint_with(big, 1, 1, image.PreprocessConfig(max_input_bytes=10))
    out = image.fingerprint_batch([big, b"junk", big], 1, [1, 2, 3])
    assert isinstance(out[1], Error) and out[0].fingerprint == out[2].fingerprint


def test_hash_index_hamming_and_jaccard(ctx):
    db = GpuIndexBackend(ctx)
    pngs = [synthetic_png(64 + 8 * i, 64)[0] for i in range(6)]
    recs = [r for r in image.fingerprint_batch(pngs, 3, list(range(10, 16)), ucfp_b200._ffi.ALGO_PHASH)]
    db.upsert(recs)
    code = image.global_hash_of(recs[2].fingerprint, recs[2].algorithm)
    hits = db.hamming_knn(3, image.ALGORITHM_PHASH, code, 3)
    assert hits[0].record_id == 12 and hits[0].score == 1.0 and len(hits) == 3
    sig = oracle.fill_u64(128, 1)
    blob = b"\x01" + bytes(7) + sig.astype("<u8").tobytes()
    near = sig.copy(); near[:28] ^= np.uint64(1)
    blob2 = b"\x01" + bytes(7) + near.astype("<u8").tobytes()
    mk = lambda rid, b: Record(3, rid, Modality.TEXT, 1, "minhash-h128", 0, b, text="x")
    db.upsert([mk(1, blob), mk(2, blob2)])
    hits = db.jaccard_knn(3, sig, 5)
    assert [h.record_id for h in hits] == [1, 2] and hits[0].score == 1.0 and hits[1].score == 100 / 128


def test_bulk_hydration_from_stored_blobs(ctx):
    """a run of stored 536-byte multi bundles / 1032-byte MinHash blobs is mirrored into HBM with one
    strided copy (ucfp_corpus_append_strided) and queried like individually upserted records."""
    from ucfp_b200 import Corpus, _ffi
    rng = np.random.default_rng(0)
    n = 5000
    words = rng.integers(0, 2**63, (n, 51), dtype=np.int64).astype(np.uint64)
    blobs = b"".join(image.pack_multihash(bytes(32), w) for w in words)
    ids = (np.arange(n, dtype=np.uint64) * np.uint64(3) + np.uint64(11))
    db = GpuIndexBackend(ctx)
    db.hydrate_fingerprints(5, image.ALGORITHM_MULTIHASH, ids, blobs)
    probe = int(words[1234, 17])                               # PHash global hash lives at offset 232
    hits = db.hamming_knn(5, image.ALGORITHM_MULTIHASH, probe, 3)
    assert hits[0].record_id == int(ids[1234]) and hits[0].score == 1.0
    oi, od = oracle.hamming_topk(np.ascontiguousarray(words[:, 17]), np.array([probe], np.uint64), 3, ids=ids)
    assert [h.record_id for h in hits] == oi[0].tolist()
    # raw ABI: MinHash payload at offset 8 of 1032-byte records, device-resident source
    import torch
    sig = oracle.fill_u64(300 * 128, 9).reshape(300, 128)
    recs = np.zeros((300, 1032), np.uint8)
    recs[:, 0] = 1
    recs[:, 8:] = sig.view(np.uint8).reshape(300, 1024)
    c = Corpus(ctx, _ffi.KIND_MINHASH128, 300)
    c.append_strided(torch.from_numpy(recs).cuda(), 1032, 8, 300)
    gi, gm = c.scan_jaccard(sig[7:8].copy(), 1)
    assert gi[0, 0] == 7 and gm[0, 0] == 128
    c.close()


def test_plain_c_client_runs_on_the_gpu(tmp_path):
    """tests/c/abi_client.c end to end: the reference's index known-answer tests, insert-or-replace / delete, the batcher and one
    image bundle, from plain C through include/ucfp_cuda.h -- the path a Rust / cgo / JNI host takes."""
    import os
    import subprocess
    from ucfp_b200 import _ffi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_client")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c", "abi_client.c"),
                    "-o", exe, "-L", libdir, "-lucfp_cuda", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
    words = [int(x, 16) for x in [l for l in out.stdout.splitlines() if l.startswith("WORDS")][0].split()[1:]]
    _, arr = synthetic_png(64, 64)
    assert words == [int(x) for x in oracle.image_multihash(arr)]
