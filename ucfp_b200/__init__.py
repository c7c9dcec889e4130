"""ucfp_b200 -- B200 (sm_100a) implementation of UCFP's data-parallel fingerprint hot path:
batched image perceptual hashing (AHash/PHash/DHash multi bundle) and the brute-force top-k scans
behind /v1/query (Hamming, MinHash-Jaccard, cosine).  Everything here sits on the C ABI of
libucfp_cuda.so (include/ucfp_cuda.h); there is no CPU fallback."""
from . import _ffi  # noqa: F401
from ._ffi import UcfpError  # noqa: F401
from .core import Error, Hit, HitSource, Modality, Query, Record  # noqa: F401
from .runtime import Batcher, Context, Corpus, Group  # noqa: F401
from . import image, sharding  # noqa: F401,E402
from .index import GpuIndexBackend  # noqa: F401,E402
from .matcher import Matcher, rrf, rrf_with_sources  # noqa: F401,E402
from . import server  # noqa: F401,E402
