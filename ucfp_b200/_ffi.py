"""ctypes binding of libucfp_cuda.so -- the same C ABI (include/ucfp_cuda.h) the reference's Rust host
binds through the `ucfp-cuda` crate (rust/ucfp-cuda/src/lib.rs).

There is no CPU fallback: if the shared library is missing, or no sm_100 GPU is present when a context is
created, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UCFP_CUDA_LIB") or os.path.join(_HERE, "libucfp_cuda.so")   # developer override: A/B runs of two builds on one box

OK, E_INVALID, E_CUDA, E_OOM, E_UNSUPPORTED, E_STATE, E_CAPACITY = 0, -1, -2, -3, -4, -5, -6
KIND_HAMMING64, KIND_MINHASH128, KIND_COSINE, KIND_MULTIHASH = 1, 2, 3, 4
ALGO_AHASH, ALGO_PHASH, ALGO_DHASH, ALGO_MULTI = 1, 2, 4, 7
PROF_HAMMING_SCAN, PROF_JACCARD_SCAN, PROF_COSINE_SCAN, PROF_IMAGE_HASH, PROF_HAMMING_TENSOR = 1, 2, 3, 4, 5
ID_NONE = 2**64 - 1

_CODE_NAMES = {E_INVALID: "UCFP_E_INVALID", E_CUDA: "UCFP_E_CUDA", E_OOM: "UCFP_E_OOM",
               E_UNSUPPORTED: "UCFP_E_UNSUPPORTED", E_STATE: "UCFP_E_STATE", E_CAPACITY: "UCFP_E_CAPACITY"}


class UcfpError(RuntimeError):
    """A negative status from the C ABI.  `.code` is the UCFP_E_* value."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{_CODE_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


class MultiHashConfig(C.Structure):
    """ucfp_multihash_config: the reference's MultiHashConfigDto (src/server/dto.rs:462-480), defaults of
    web/src/lib/docs/api-reference-image.md:51-62."""
    _fields_ = [("ahash_weight", C.c_float), ("phash_weight", C.c_float), ("dhash_weight", C.c_float),
                ("global_weight", C.c_float), ("block_weight", C.c_float), ("block_distance_threshold", C.c_uint32)]

    @classmethod
    def of(cls, cfg=None):
        d = {"ahash_weight": 0.1, "phash_weight": 0.4, "dhash_weight": 0.3, "global_weight": 0.1, "block_weight": 0.1,
             "block_distance_threshold": 12}
        d.update(cfg or {})
        return cls(d["ahash_weight"], d["phash_weight"], d["dhash_weight"], d["global_weight"], d["block_weight"],
                   int(d["block_distance_threshold"]))


class ImageDesc(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("stride", C.c_uint64)]


# every symbol include/ucfp_cuda.h declares: name -> (restype, argtypes)
_vp, _sz, _u64, _u32, _int = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int
PROTOTYPES = {
    "ucfp_init": (_int, [_int, C.POINTER(_vp)]),
    "ucfp_destroy": (None, [_vp]),
    "ucfp_ctx_set_stream": (_int, [_vp, _vp]),
    "ucfp_ctx_reset_stream": (_int, [_vp]),
    "ucfp_ctx_synchronize": (_int, [_vp]),
    "ucfp_abi_version": (_int, []),
    "ucfp_last_error": (C.c_char_p, []),
    "ucfp_ctx_kernel_launches": (_u64, [_vp]),
    "ucfp_ctx_profile_begin": (_int, [_vp]),
    "ucfp_ctx_profile_end": (_int, [_vp, _int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_u64)]),
    "ucfp_ctx_profile_read": (_int, [_vp, _int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_u64)]),
    "ucfp_image_hash_batch": (_int, [_vp, C.POINTER(ImageDesc), _sz, _u32, _vp, _vp]),
    "ucfp_image_hash_uniform": (_int, [_vp, _vp, _sz, _u32, _u32, _u64, _u64, _u32, _vp]),
    "ucfp_image_hash_jpeg_batch": (_int, [_vp, C.POINTER(_vp), C.POINTER(_sz), _sz, _u32, _vp, _vp, _vp, _vp, _sz]),
    "ucfp_corpus_create": (_int, [_vp, _int, _u32, _u64, C.POINTER(_vp)]),
    "ucfp_corpus_destroy": (None, [_vp]),
    "ucfp_corpus_append": (_int, [_vp, _vp, _vp, _u64]),
    "ucfp_corpus_append_strided": (_int, [_vp, _vp, _vp, _u64, _u64, _u64]),
    "ucfp_corpus_set_id_base": (_int, [_vp, _u64]),
    "ucfp_corpus_clear": (_int, [_vp]),
    "ucfp_corpus_size": (_u64, [_vp]),
    "ucfp_corpus_append_synthetic": (_int, [_vp, _u64, _u64, _u64]),
    "ucfp_corpus_device_rows": (_vp, [_vp]),
    "ucfp_corpus_refresh": (_int, [_vp]),
    "ucfp_corpus_upsert": (_int, [_vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    "ucfp_corpus_delete": (_int, [_vp, _vp, _u64, C.POINTER(_u64)]),
    "ucfp_corpus_reserve": (_int, [_vp, _u64]),
    "ucfp_corpus_capacity": (_u64, [_vp]),
    "ucfp_scan_hamming": (_int, [_vp, _vp, _sz, _sz, _vp, _vp]),
    "ucfp_scan_jaccard": (_int, [_vp, _vp, _sz, _sz, _vp, _vp]),
    "ucfp_scan_cosine": (_int, [_vp, _vp, _sz, _sz, _vp, _vp]),
    "ucfp_ctx_last_scan_fallbacks": (_int, [_vp, C.POINTER(_u64)]),
    "ucfp_ctx_last_scan_stats": (_int, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "ucfp_ctx_last_scan_exact_selects": (_int, [_vp, C.POINTER(_u64)]),
    "ucfp_scan_multihash": (_int, [_vp, _vp, _sz, _sz, _sz, C.POINTER(MultiHashConfig), _vp, _vp]),
    "ucfp_multihash_compare": (_int, [_vp, _vp, _vp, _sz, C.POINTER(MultiHashConfig), _vp]),
    "ucfp_batcher_create": (_int, [_vp, _u32, _u32, C.POINTER(_vp)]),
    "ucfp_batcher_destroy": (None, [_vp]),
    "ucfp_batcher_query": (_int, [_vp, _vp, _sz, _vp, _vp]),
    "ucfp_batcher_stats": (_int, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "ucfp_group_create": (_int, [C.POINTER(_int), _int, C.POINTER(_vp)]),
    "ucfp_group_unique_id": (_int, [_vp]),
    "ucfp_group_join": (_int, [_vp, _vp, _int, _int, C.POINTER(_vp)]),
    "ucfp_group_destroy": (None, [_vp]),
    "ucfp_group_local_size": (_int, [_vp]),
    "ucfp_group_world_size": (_int, [_vp]),
    "ucfp_group_ctx": (_vp, [_vp, _int]),
    "ucfp_group_scan_hamming": (_int, [_vp, C.POINTER(_vp), _vp, _sz, _sz, _vp, _vp]),
    "ucfp_group_scan_jaccard": (_int, [_vp, C.POINTER(_vp), _vp, _sz, _sz, _vp, _vp]),
    "ucfp_group_scan_cosine": (_int, [_vp, C.POINTER(_vp), _vp, _sz, _sz, _vp, _vp]),
    "ucfp_merge_topk_u32": (_int, [_vp, _vp, _vp, _sz, _sz, _sz, _int, _vp, _vp]),
    "ucfp_merge_topk_f32": (_int, [_vp, _vp, _vp, _sz, _sz, _sz, _vp, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads libucfp_cuda.so.  Raises if it has not been built -- there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m ucfp_b200.build` (nvcc, sm_100a). "
                "ucfp_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise UcfpError(rc, lib().ucfp_last_error().decode("utf-8", "replace"))
