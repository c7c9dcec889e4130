#include "common.cuh"
namespace ucfp {
int cosine_on_append(ucfp_corpus *, uint64_t, uint64_t) { return UCFP_OK; }
int cosine_scan(ucfp_corpus *, const float *, size_t, size_t, uint64_t *, float *) {
    set_error("cosine scan not built yet"); return UCFP_E_UNSUPPORTED;
}
}
