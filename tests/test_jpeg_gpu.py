"""batch ingest from ENCODED bytes -- JPEG decoded on the device (nvJPEG) and hashed by the same call.
Parity: the hashes must equal the oracle run over the very pixels the device decoded (bit-exact, as everywhere), and those
pixels must agree with the host decoder the reference side would use (Pillow / libjpeg-turbo here) to within the usual
IDCT / chroma-upsampling tolerance -- JPEG decoders are not bit-identical to each other and nobody claims they are."""
import io

import numpy as np
import pytest

import oracle
from ucfp_b200 import _ffi

pytestmark = pytest.mark.gpu


def _jpeg(arr, quality=90, subsampling=0, progressive=False):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="JPEG", quality=quality, subsampling=subsampling, progressive=progressive)
    return buf.getvalue()


def _scene(w, h, seed):
    """smooth synthetic photo-like content (JPEG on noise is meaningless)"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        for _ in range(6):
            fx, fy, ph = rng.uniform(0.002, 0.05, 2).tolist() + [rng.uniform(0, 6.28)]
            img[..., c] += rng.uniform(20, 60) * np.sin(fx * x + fy * y + ph)
    img += 128
    return np.clip(img, 0, 255).astype(np.uint8)


def test_jpeg_batch_decodes_on_the_device_and_hashes_bit_exactly(ctx):
    from PIL import Image
    shapes = [(256, 256), (640, 480), (1024, 1024), (333, 517), (64, 48), (1920, 1080)]
    blobs = [_jpeg(_scene(w, h, i), quality=q, subsampling=s) for i, ((w, h), q, s) in
             enumerate(zip(shapes, (90, 75, 95, 85, 90, 80), (0, 2, 1, 2, 0, 2)))]
    words, status, dims, pixels = ctx.image_hash_jpeg_batch(blobs, want_pixels=True)
    assert (status == 0).all(), status
    for i, (w, h) in enumerate(shapes):
        assert tuple(dims[i]) == (w, h) and pixels[i].shape == (h, w, 3)
        np.testing.assert_array_equal(words[i], oracle.image_multihash(pixels[i]))          # bit-exact on the decoded pixels
        host = np.asarray(Image.open(io.BytesIO(blobs[i])).convert("RGB"))
        diff = np.abs(host.astype(np.int16) - pixels[i].astype(np.int16))
        # decoders agree to IDCT rounding where chroma is not subsampled, and to the chroma-upsampling filter (libjpeg-turbo's
        # "fancy" triangle filter vs nvJPEG's) where it is: a few grey levels on average, more at sharp colour edges
        assert diff.mean() < 3.0 and np.percentile(diff, 99) <= 16, (i, diff.max(), diff.mean())
    # without the pixel read-back the hashes are the same
    words2, status2, _ = ctx.image_hash_jpeg_batch(blobs)
    np.testing.assert_array_equal(words2, words)
    # single-algorithm mask
    w3, s3, _ = ctx.image_hash_jpeg_batch(blobs[:2], _ffi.ALGO_PHASH)
    assert (s3 == 0).all() and (w3[:, 17:34] == words[:2, 17:34]).all() and (w3[:, :17] == 0).all() and (w3[:, 34:] == 0).all()


def test_bad_and_foreign_inputs_get_per_image_status(ctx):
    from PIL import Image
    good = _jpeg(_scene(128, 96, 1))
    png = io.BytesIO(); Image.fromarray(_scene(64, 64, 2)).save(png, format="PNG")
    truncated = good[: len(good) // 3]
    tiny = _jpeg(_scene(8, 8, 3))[:0] + _jpeg(np.zeros((2, 2, 3), np.uint8))
    words, status, dims = ctx.image_hash_jpeg_batch([good, png.getvalue(), b"", truncated, tiny, good])
    assert status[0] == 0 and status[5] == 0 and (words[0] == words[5]).all() and words[0].any()
    assert status[1] == _ffi.E_UNSUPPORTED                      # PNG: the host decodes it and calls ucfp_image_hash_batch
    assert status[2] == _ffi.E_INVALID and status[4] in (_ffi.E_UNSUPPORTED, _ffi.E_INVALID)
    assert status[3] != 0 or not words[3].any() or True         # a truncated stream may decode to garbage or fail; it must not fail the batch
    assert (words[1] == 0).all() and (words[2] == 0).all()
    # progressive JPEG: decoded, or handed back to the host -- never a batch failure
    prog = _jpeg(_scene(200, 120, 4), progressive=True)
    w2, s2, _, px = ctx.image_hash_jpeg_batch([prog, good], want_pixels=True)
    assert s2[1] == 0 and s2[0] in (0, _ffi.E_UNSUPPORTED)
    if s2[0] == 0:
        np.testing.assert_array_equal(w2[0], oracle.image_multihash(px[0]))


def test_ingest_throughput_from_encoded_bytes_is_reported(ctx):
    """Not a gate on a number: prints images/s from encoded bytes for a 256-image 1024x1024 batch (decode + hash on the device)."""
    import time
    blobs = [_jpeg(_scene(1024, 1024, 100 + i % 8), quality=85, subsampling=2) for i in range(8)] * 32
    ctx.image_hash_jpeg_batch(blobs[:16])
    t0 = time.perf_counter()
    words, status, _ = ctx.image_hash_jpeg_batch(blobs)
    dt = time.perf_counter() - t0
    assert (status == 0).all() and (words[0] == words[8]).all()
    mb = sum(len(b) for b in blobs) / 1e6
    print(f"\njpeg ingest: {len(blobs)} x 1024x1024 ({mb:.0f} MB encoded) in {dt * 1e3:.1f} ms = {len(blobs) / dt:.0f} images/s from encoded bytes")
