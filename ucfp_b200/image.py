"""Image fingerprinting -- mirror of the reference's src/modality/image.rs, hashing on the GPU.

Same names, argument meaning and error behaviour: `fingerprint`, `fingerprint_with`,
`fingerprint_multi_with`, `fingerprint_phash|dhash|ahash` take ENCODED image bytes plus tenant/record ids and
return a `Record` whose `fingerprint` blob has imgfprint's layout:

    ImageFingerprint      168 B = exact[32] | global_hash u64 LE | block_hashes [u64; 16] LE
                          (web/src/lib/components/charts/ImageHashView.svelte:2-5)
    MultiHashFingerprint  536 B = exact[32] | ahash(168) | phash(168) | dhash(168) at offsets 0/32/200/368
                          (web/src/lib/components/charts/AlgorithmView.svelte:30-37, server/tests.rs:1206)

What stays on the host, as in the reference: decoding (Pillow here, the `image` crate there), the
PreprocessConfig guards of build_image_preprocess (src/server/handlers.rs:307-319), and BLAKE3 `exact`.
What moves to the GPU: everything after decode (ucfp_image_hash_batch).  Errors raise
`Error("Modality", ...)`, the analogue of Error::Modality (src/modality/image.rs:70).

`fingerprint_batch` is the batched entry point the reference lacks (one image per HTTP request there).
"""
from __future__ import annotations

import io
import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from .core import Error, Modality, Record
from .runtime import Context

# The reference's tags (src/modality/image.rs:38-46) name imgfprint's hash definitions.  Bit parity of this repo's hashes
# with imgfprint 0.4.1 is UNPINNED (docs/HASH_SPEC.md, tests/test_imgfprint_parity.py), and an index keys its Hamming
# corpora by algorithm tag: stamping spec-v1 hashes with imgfprint's tags would let records hashed here and records
# hashed by the reference share one corpus and return meaningless distances.  Until the golden vectors exist and the
# parity test passes, records produced here carry their own tags; flip the flag below (and only then) to become a
# byte-for-byte drop-in.  Records hydrated from the reference's store keep the tags they were written with.
IMGFPRINT_PARITY_VERIFIED = False
REFERENCE_ALGORITHM_MULTIHASH = "imgfprint-multihash-v1"  # src/modality/image.rs:38, :40
REFERENCE_ALGORITHM_PHASH = "imgfprint-phash-v1"          # :42
REFERENCE_ALGORITHM_DHASH = "imgfprint-dhash-v1"          # :44
REFERENCE_ALGORITHM_AHASH = "imgfprint-ahash-v1"          # :46
_PREFIX = "imgfprint" if IMGFPRINT_PARITY_VERIFIED else "ucfp-b200"
ALGORITHM = f"{_PREFIX}-multihash-v1"
ALGORITHM_MULTIHASH = f"{_PREFIX}-multihash-v1"
ALGORITHM_PHASH = f"{_PREFIX}-phash-v1"
ALGORITHM_DHASH = f"{_PREFIX}-dhash-v1"
ALGORITHM_AHASH = f"{_PREFIX}-ahash-v1"
MULTIHASH_TAGS = (ALGORITHM_MULTIHASH, REFERENCE_ALGORITHM_MULTIHASH)   # 536-byte bundles
SINGLE_HASH_TAGS = (ALGORITHM_PHASH, ALGORITHM_DHASH, ALGORITHM_AHASH, REFERENCE_ALGORITHM_PHASH, REFERENCE_ALGORITHM_DHASH,
                    REFERENCE_ALGORITHM_AHASH)                            # 168-byte ImageFingerprints
FORMAT_VERSION = 1                               # imgfprint::FORMAT_VERSION (web/src/lib/docs/api-reference-image.md:103)

_TAG = {_ffi.ALGO_MULTI: ALGORITHM_MULTIHASH, _ffi.ALGO_PHASH: ALGORITHM_PHASH, _ffi.ALGO_DHASH: ALGORITHM_DHASH,
        _ffi.ALGO_AHASH: ALGORITHM_AHASH}


@dataclass
class PreprocessConfig:
    """imgfprint::PreprocessConfig; defaults from src/server/algorithms_manifest.rs:446-469."""
    max_input_bytes: int = 50 * 1024 * 1024
    max_dimension: int = 8192
    min_dimension: int = 32


def decode_rgb(data: bytes, preprocess: PreprocessConfig) -> np.ndarray:
    """Host-side decode + guards.  Returns (h, w, 3) u8."""
    if len(data) == 0:
        raise Error("Modality", "empty input")
    if len(data) > preprocess.max_input_bytes:
        raise Error("Modality", f"input of {len(data)} bytes exceeds max_input_bytes {preprocess.max_input_bytes}")
    try:
        from PIL import Image, ImageOps
        im = Image.open(io.BytesIO(data))
        im = ImageOps.exif_transpose(im)
        rgb = np.asarray(im.convert("RGB"), dtype=np.uint8)
    except Error:
        raise
    except Exception as e:  # garbage bytes -> 400 in the reference (server/tests.rs:267-280)
        raise Error("Modality", f"image decode: {e}") from e
    h, w = rgb.shape[:2]
    if max(w, h) > preprocess.max_dimension:
        raise Error("Modality", f"image {w}x{h} exceeds max_dimension {preprocess.max_dimension}")
    if min(w, h) < preprocess.min_dimension:
        raise Error("Modality", f"image {w}x{h} is below min_dimension {preprocess.min_dimension}")
    return np.ascontiguousarray(rgb)


def exact_hash(data: bytes) -> bytes:
    """`exact: [u8; 32]` = BLAKE3-256 of the encoded input (AlgorithmView.svelte:30-33); host side."""
    import blake3
    return blake3.blake3(data).digest()


def pack_image_fingerprint(exact: bytes, hash17: Sequence[int]) -> bytes:
    """168-byte ImageFingerprint."""
    assert len(exact) == 32 and len(hash17) == 17
    return exact + struct.pack("<17Q", *[int(x) for x in hash17])


def pack_multihash(exact: bytes, words51: Sequence[int]) -> bytes:
    """536-byte MultiHashFingerprint from the 51 words of ucfp_image_hashes (ahash | phash | dhash)."""
    assert len(words51) == 51
    w = [int(x) for x in words51]
    return exact + b"".join(pack_image_fingerprint(exact, w[17 * a: 17 * a + 17]) for a in range(3))


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _record(tenant_id: int, record_id: int, tag: str, blob: bytes) -> Record:
    # src/modality/image.rs:72-87 / :181-193
    return Record(tenant_id=tenant_id, record_id=record_id, modality=Modality.IMAGE, format_version=FORMAT_VERSION,
                  algorithm=tag, config_hash=0, fingerprint=blob, embedding=None, model_id=None, metadata=b"", text=None)


def fingerprint_batch(items: Sequence[bytes], tenant_id: int, record_ids: Sequence[int], algo_mask: int = _ffi.ALGO_MULTI,
                      preprocess: Optional[PreprocessConfig] = None, ctx: Optional[Context] = None) -> List:
    """Batched ingest: decode on the host, hash the whole batch in one GPU call.  Returns a list with a
    Record or an Error per input (one bad image does not fail the batch)."""
    preprocess = preprocess or PreprocessConfig()
    ctx = ctx or default_context()
    decoded, results = [], [None] * len(items)
    for i, data in enumerate(items):
        try:
            decoded.append(decode_rgb(data, preprocess))
        except Error as e:
            decoded.append(None)
            results[i] = e
    ok = [i for i, d in enumerate(decoded) if d is not None]
    if ok:
        words, status = ctx.image_hash_batch([decoded[i] for i in ok], algo_mask)
        for j, i in enumerate(ok):
            if status[j] != 0:
                results[i] = Error("Modality", f"image hash failed with status {int(status[j])}")
                continue
            exact = exact_hash(items[i])
            if algo_mask == _ffi.ALGO_MULTI:
                blob = pack_multihash(exact, words[j])
            else:
                a = {_ffi.ALGO_AHASH: 0, _ffi.ALGO_PHASH: 1, _ffi.ALGO_DHASH: 2}[algo_mask]
                blob = pack_image_fingerprint(exact, words[j][17 * a: 17 * a + 17])
            results[i] = _record(tenant_id, record_ids[i], _TAG[algo_mask], blob)
    return results


def _single(data: bytes, tenant_id: int, record_id: int, algo_mask: int, preprocess: Optional[PreprocessConfig]) -> Record:
    r = fingerprint_batch([data], tenant_id, [record_id], algo_mask, preprocess)[0]
    if isinstance(r, Error):
        raise r
    return r


def fingerprint(data: bytes, tenant_id: int, record_id: int) -> Record:
    """src/modality/image.rs:56-58."""
    return fingerprint_with(data, tenant_id, record_id, PreprocessConfig())


def fingerprint_with(data: bytes, tenant_id: int, record_id: int, preprocess: PreprocessConfig) -> Record:
    """src/modality/image.rs:62-88: the multi-hash bundle."""
    return _single(data, tenant_id, record_id, _ffi.ALGO_MULTI, preprocess)


def fingerprint_multi_with(data: bytes, preprocess: PreprocessConfig, _multi_cfg, tenant_id: int, record_id: int) -> Record:
    """src/modality/image.rs:96-104: the compare-time config does not affect the stored bytes."""
    return fingerprint_with(data, tenant_id, record_id, preprocess)


def fingerprint_phash(data: bytes, preprocess: PreprocessConfig, tenant_id: int, record_id: int) -> Record:
    """src/modality/image.rs:112-126."""
    return _single(data, tenant_id, record_id, _ffi.ALGO_PHASH, preprocess)


def fingerprint_dhash(data: bytes, preprocess: PreprocessConfig, tenant_id: int, record_id: int) -> Record:
    """src/modality/image.rs:130-144."""
    return _single(data, tenant_id, record_id, _ffi.ALGO_DHASH, preprocess)


def fingerprint_ahash(data: bytes, preprocess: PreprocessConfig, tenant_id: int, record_id: int) -> Record:
    """src/modality/image.rs:148-162."""
    return _single(data, tenant_id, record_id, _ffi.ALGO_AHASH, preprocess)


# ---- field extraction used when hydrating the scan corpora from stored records -------
def global_hash_of(fingerprint: bytes, algorithm: str) -> int:
    """The u64 the Hamming scan indexes: global_hash @32 of a 168-byte ImageFingerprint, or the PHash
    global hash @232 of a 536-byte multi bundle."""
    if algorithm in MULTIHASH_TAGS:
        if len(fingerprint) != 536:
            raise Error("Incompatible", f"multihash bundle must be 536 bytes, got {len(fingerprint)}")
        return struct.unpack_from("<Q", fingerprint, 232)[0]
    if len(fingerprint) != 168:
        raise Error("Incompatible", f"image fingerprint must be 168 bytes, got {len(fingerprint)}")
    return struct.unpack_from("<Q", fingerprint, 32)[0]


def minhash_payload_of(fingerprint: bytes) -> np.ndarray:
    """The 128 u64 slots the Jaccard scan indexes: payload @8 of txtfp's 1032-byte MinHashSig<128>
    {schema: u16 = 1, _pad: [u8; 6], hashes: [u64; 128]} (src/modality/text.rs:200-204, server/tests.rs:1153-1162)."""
    if len(fingerprint) != 1032:
        raise Error("Incompatible", f"MinHashSig<128> must be 1032 bytes, got {len(fingerprint)}")
    if fingerprint[:8] != b"\x01\x00\x00\x00\x00\x00\x00\x00":
        raise Error("Incompatible", "unsupported MinHashSig schema header")
    return np.frombuffer(fingerprint, dtype="<u8", count=128, offset=8).copy()


def bundle_words_of(fingerprint: bytes) -> np.ndarray:
    """The 51 hash words (ahash[17] | phash[17] | dhash[17]) of a 536-byte multi bundle: what a MULTIHASH corpus row holds."""
    if len(fingerprint) != 536:
        raise Error("Incompatible", f"multihash bundle must be 536 bytes, got {len(fingerprint)}")
    return np.concatenate([np.frombuffer(fingerprint, dtype="<u8", count=17, offset=64 + 168 * a) for a in range(3)])
