"""Parity of the CUDA image multi-hash (AHash + PHash + DHash, 1 global + 16 block hashes each) against the
CPU oracle, through the C ABI.  Hash bits must be bit-exact (docs/HASH_SPEC.md).  The reference's own tests
pin no hash bit (server/tests.rs:239-263,456-532,1167-1208 assert tags and the 536-byte size), so parity with
imgfprint itself is unpinned; see tests/test_imgfprint_parity.py."""
import numpy as np
import pytest

import oracle
from ucfp_b200 import _ffi

pytestmark = pytest.mark.gpu


def ramp(w, h):
    """synthetic_png of the reference (src/server/tests.rs:227-235, benches/end_to_end.rs:77-85), decoded."""
    y, x = np.mgrid[0:h, 0:w]
    return np.stack([x % 256, y % 256, np.full_like(x, 128)], -1).astype(np.uint8)


def noise(w, h, seed):
    return oracle.fill_u64((w * h * 3 + 7) // 8, seed).view(np.uint8)[: w * h * 3].reshape(h, w, 3).copy()


def photo_like(w, h, seed):
    """smooth blobs + edges + mild noise: hashes that are not coin flips"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for _ in range(6):
        cx, cy, s = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(0.05, 0.4) * max(w, h)
        col = rng.uniform(0, 255, 3)
        img += np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * s * s))[..., None] * col
    img[: h // 3, w // 2:] *= 0.5
    img += rng.normal(0, 4, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def check(ctx, images, algo_mask=_ffi.ALGO_MULTI):
    got, status = ctx.image_hash_batch(images, algo_mask)
    assert (status == 0).all(), status
    for i, im in enumerate(images):
        want = oracle.image_multihash(im)
        if not algo_mask & _ffi.ALGO_AHASH:
            want[0:17] = 0
        if not algo_mask & _ffi.ALGO_PHASH:
            want[17:34] = 0
        if not algo_mask & _ffi.ALGO_DHASH:
            want[34:51] = 0
        bad = np.nonzero(got[i] != want)[0]
        assert bad.size == 0, f"image {i} {im.shape}: words {bad[:8]} differ: got {got[i][bad[:4]]} want {want[bad[:4]]}"


def test_reference_ramp_images(ctx):
    check(ctx, [ramp(256, 256), ramp(64, 64), ramp(1024, 1024)])


@pytest.mark.parametrize("w,h", [(256, 256), (1024, 1024), (640, 480), (1000, 700), (1920, 1080), (513, 129), (128, 128),
                                 (2048, 1536), (4096, 512),
                                 # two columns per thread (even widths 128..512): bulk-staged, halfword and ragged variants; odd widths stay on one
                                 (130, 200), (132, 140), (384, 216), (510, 300), (512, 512), (200, 200), (257, 300), (511, 64)])
def test_stream_kernel_shapes(ctx, w, h):
    check(ctx, [noise(w, h, 1), photo_like(w, h, 2)])


@pytest.mark.parametrize("w,h", [(4, 4), (5, 7), (31, 31), (32, 32), (37, 53), (64, 64), (100, 300), (127, 500), (4100, 40),
                                 (4500, 300), (33, 2000)])
def test_generic_kernel_shapes(ctx, w, h):
    check(ctx, [noise(w, h, 3), photo_like(w, h, 4)])


def test_mixed_batch_and_bad_images(ctx):
    imgs = [noise(256, 256, 10), photo_like(300, 200, 11), noise(1024, 1024, 12), noise(256, 256, 13), ramp(640, 360)]
    check(ctx, imgs)
    got, status = ctx.image_hash_batch([imgs[0], None, np.zeros((2, 2, 3), np.uint8), imgs[1]])
    assert status.tolist() == [0, _ffi.E_INVALID, _ffi.E_INVALID, 0]
    assert (got[1] == 0).all() and (got[2] == 0).all()
    np.testing.assert_array_equal(got[0], oracle.image_multihash(imgs[0]))
    np.testing.assert_array_equal(got[3], oracle.image_multihash(imgs[1]))


@pytest.mark.parametrize("mask", [_ffi.ALGO_AHASH, _ffi.ALGO_PHASH, _ffi.ALGO_DHASH, _ffi.ALGO_PHASH | _ffi.ALGO_DHASH])
def test_single_algorithm_masks(ctx, mask):
    check(ctx, [photo_like(256, 256, 5), noise(640, 480, 6)], mask)


def test_uniform_device_batch_matches_oracle(ctx):
    """Batch-ingest layout: n equally sized images resident in HBM, outputs in HBM."""
    import torch
    n, w, h = 64, 256, 256
    host = np.stack([noise(w, h, 100 + i) if i % 2 else photo_like(w, h, 100 + i) for i in range(n)])
    dev = torch.from_numpy(host).cuda()
    out = ctx.image_hash_uniform(dev, n, w, h)
    torch.cuda.synchronize()
    want = oracle.image_multihash_batch(host, threads=oracle.host_threads())
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint64), want)


def test_config1_10k_images_then_hamming(ctx):
    """BASELINE config 1 (the reference's CPU-runnable case): multi bundle on 10 000 synthetic 256x256 RGB
    images (image 0 = the reference ramp), then Hamming top-10 over their PHash codes."""
    import torch
    from ucfp_b200 import Corpus
    n, w, h = 10_000, 256, 256
    words = w * h * 3 // 8
    host = oracle.fill_u64(n * words, 0x1316).view(np.uint8).reshape(n, h, w, 3).copy()
    host[0] = ramp(w, h)
    out = ctx.image_hash_uniform(torch.from_numpy(host).cuda(), n, w, h).cpu().numpy().view(np.uint64)
    want = oracle.image_multihash_batch(host, threads=oracle.host_threads())
    np.testing.assert_array_equal(out, want)
    phash = np.ascontiguousarray(out[:, 17])
    corpus = Corpus(ctx, _ffi.KIND_HAMMING64, n)
    corpus.append(phash)
    ids, dist = corpus.scan_hamming(phash[:1024].copy(), 10)
    oi, od = oracle.hamming_topk(want[:, 17].copy(), want[:1024, 17].copy(), 10, threads=oracle.host_threads())
    np.testing.assert_array_equal(ids, oi)
    np.testing.assert_array_equal(dist, od)
    assert (dist[:, 0] == 0).all()  # every query finds itself
    corpus.close()
