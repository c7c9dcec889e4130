#include "common.cuh"
namespace ucfp {
int jaccard_on_append(ucfp_corpus *, uint64_t, uint64_t) { return UCFP_OK; }
int jaccard_scan(ucfp_corpus *, const uint64_t *, size_t, size_t, uint64_t *, uint32_t *) {
    set_error("jaccard scan not built yet"); return UCFP_E_UNSUPPORTED;
}
}
