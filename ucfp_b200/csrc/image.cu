#include "common.cuh"
namespace ucfp {
int image_hash_batch(ucfp_ctx *, const ucfp_image_desc *, size_t, uint32_t, ucfp_image_hashes *, int32_t *) {
    set_error("image hashing not built yet"); return UCFP_E_UNSUPPORTED;
}
}
