"""Regenerates tests/golden/oracle_golden.json from the CPU oracle.

These vectors pin the ORACLE (docs/HASH_SPEC.md) against regressions and travel to the GPU box, where
/root/reference does not exist.  They are not reference outputs: the reference cannot be built here (no
cargo; imgfprint is not vendored) and its own tests hold no image-hash, Hamming or Jaccard vectors
(SURVEY F2, F3, F5).  The reference-anchored checks live in tests/test_oracle.py (cosine known answers,
MinHash layout header).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def ramp(w, h):
    y, x = np.mgrid[0:h, 0:w]
    return np.stack([x % 256, y % 256, np.full_like(x, 128)], -1).astype(np.uint8)


def noise(w, h, seed):
    return oracle.fill_u64((w * h * 3 + 7) // 8, seed).view(np.uint8)[: w * h * 3].reshape(h, w, 3).copy()


def main():
    g = {"spec": "docs/HASH_SPEC.md v1", "images": [], "hamming": {}, "jaccard": {}, "prng": {}}
    for name, img in [("ramp256", ramp(256, 256)), ("ramp64", ramp(64, 64)), ("ramp300x200", ramp(300, 200)),
                      ("noise256_s1", noise(256, 256, 1)), ("noise1024_s2", noise(1024, 1024, 2)), ("noise37x53_s3", noise(37, 53, 3)),
                      ("noise640x480_s4", noise(640, 480, 4))]:
        g["images"].append({"name": name, "words": [f"{int(v):016x}" for v in oracle.image_multihash(img)]})
    g["prng"] = {"seed": 0xC0DE, "first8": [f"{int(v):016x}" for v in oracle.fill_u64(8, 0xC0DE)]}
    codes, q = oracle.fill_u64(50_000, 0xC0DE), oracle.fill_u64(4, 0xBEEF)
    ids, d = oracle.hamming_topk(codes, q, 10)
    g["hamming"] = {"n": 50_000, "seed_codes": 0xC0DE, "seed_queries": 0xBEEF, "k": 10, "ids": ids.tolist(), "dist": d.tolist()}
    sig = oracle.fill_u64(2000 * 128, 7).reshape(2000, 128)
    qs = oracle.fill_u64(2 * 128, 8).reshape(2, 128)
    sig[1234, :100] = qs[0, :100]
    sig[77, 5:70] = qs[1, 5:70]
    sig[78, 5:70] = qs[1, 5:70]
    ids, m = oracle.jaccard_topk(sig, qs, 5)
    g["jaccard"] = {"n": 2000, "seed_sigs": 7, "seed_queries": 8, "k": 5, "ids": ids.tolist(), "matches": m.tolist()}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w") as f:
        json.dump(g, f, indent=1)


if __name__ == "__main__":
    main()
