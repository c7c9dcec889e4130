// merge.cu -- deterministic merge of per-shard top-k lists (the step after the NCCL all-gather of
// per-rank candidates; replaces rayon's reduce of per-thread buffers, src/index/embedded/mod.rs:335-342).
// One CTA per query; bitonic sort of parts*k (id, key) pairs in shared memory under the total order
// (key best-first, record_id asc), sentinels (id == UCFP_ID_NONE) last.
#include <math.h>

#include "common.cuh"

namespace ucfp {
namespace {

template <typename K> struct Order;
template <> struct Order<uint32_t> {
    int descending;
    __device__ bool before(uint32_t ka, uint64_t ia, uint32_t kb, uint64_t ib) const {
        if (ia == UINT64_MAX || ib == UINT64_MAX) return ia != UINT64_MAX && ib == UINT64_MAX;
        if (ka != kb) return descending ? ka > kb : ka < kb;
        return ia < ib;
    }
};
template <> struct Order<float> {
    int descending;
    __device__ bool before(float ka, uint64_t ia, float kb, uint64_t ib) const {
        if (ia == UINT64_MAX || ib == UINT64_MAX) return ia != UINT64_MAX && ib == UINT64_MAX;
        if (ka != kb) return ka > kb;
        return ia < ib;
    }
};

template <typename K>
__global__ void merge_topk_kernel(const uint64_t *__restrict__ ids_in, const K *__restrict__ keys_in, uint32_t parts,
                                  uint32_t nq, uint32_t k, Order<K> ord, K sentinel, uint64_t *ids_out, K *keys_out) {
    extern __shared__ unsigned char raw[];
    const uint32_t q = blockIdx.x;
    const uint32_t n = parts * k;
    uint32_t P = 1;
    while (P < n) P <<= 1;
    uint64_t *s_id = reinterpret_cast<uint64_t *>(raw);
    K *s_key = reinterpret_cast<K *>(s_id + P);
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        uint64_t id = UINT64_MAX; K key = sentinel;
        if (i < n) {
            uint32_t p = i / k, j = i % k;
            size_t src = ((size_t)p * nq + q) * k + j;
            if (keys_in) { id = ids_in[src]; key = keys_in[src]; }
            else {   // packed 16-byte records {u64 id; 32-bit key; pad} as all-gathered by a group scan
                const ulonglong2 rec = reinterpret_cast<const ulonglong2 *>(ids_in)[src];
                const uint32_t bits = (uint32_t)rec.y;
                id = rec.x; key = *reinterpret_cast<const K *>(&bits);
            }
        }
        s_id[i] = id; s_key[i] = key;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= P; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                uint32_t i = 2 * t - (t & (stride - 1)), j = i + stride;
                bool up = ((i & size) == 0);
                uint64_t ii = s_id[i], ij = s_id[j]; K ki = s_key[i], kj = s_key[j];
                bool swap = up ? ord.before(kj, ij, ki, ii) : ord.before(ki, ii, kj, ij);
                if (swap) { s_id[i] = ij; s_key[i] = kj; s_id[j] = ii; s_key[j] = ki; }
            }
            __syncthreads();
        }
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t id = s_id[i];
        ids_out[(size_t)q * k + i] = id;
        keys_out[(size_t)q * k + i] = id == UINT64_MAX ? sentinel : s_key[i];
    }
}

template <typename K>
int merge_impl(ucfp_lane *ctx, const uint64_t *ids_in, const K *keys_in, size_t parts, size_t nq, size_t k, int descending,
               K sentinel, uint64_t *ids_out, K *keys_out) {
    if (nq == 0 || k == 0) return UCFP_OK;
    UCFP_REQUIRE(parts >= 1 && parts * k <= 16384, UCFP_E_UNSUPPORTED, "merge supports parts*k <= 16384 (got %zu)", parts * k);
    size_t P = 1;
    while (P < parts * k) P <<= 1;
    size_t smem = P * (sizeof(uint64_t) + sizeof(K));
    auto kern = merge_topk_kernel<K>;   // dynamic shared memory opted in by merge_device_init
    kern<<<(unsigned)nq, 256, smem, ctx->stream>>>(ids_in, keys_in, (uint32_t)parts, (uint32_t)nq, (uint32_t)k,
                                                    Order<K>{descending}, sentinel, ids_out, keys_out);
    count_launch(ctx);
    return check_launch("merge_topk");
}

}  // namespace

int merge_device_init(ucfp_ctx *) {
    UCFP_CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(16384 * 12)));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(16384 * 12)));
    return UCFP_OK;
}

int merge_u32(ucfp_lane *ctx, const uint64_t *ids_in, const uint32_t *keys_in, size_t parts, size_t nq, size_t k,
              int descending, uint64_t *ids_out, uint32_t *keys_out) {
    return merge_impl<uint32_t>(ctx, ids_in, keys_in, parts, nq, k, descending, UINT32_MAX, ids_out, keys_out);
}

int merge_f32(ucfp_lane *ctx, const uint64_t *ids_in, const float *keys_in, size_t parts, size_t nq, size_t k,
              uint64_t *ids_out, float *keys_out) {
    return merge_impl<float>(ctx, ids_in, keys_in, parts, nq, k, 1, -INFINITY, ids_out, keys_out);
}

// ---- packed (id, key) records: what a group scan exchanges (one all-gather instead of two) ------------------------
namespace {
__global__ void pack_topk_kernel(const uint64_t *__restrict__ ids, const uint32_t *__restrict__ keys, size_t n, ulonglong2 *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_ulonglong2(ids[i], (unsigned long long)keys[i]);
}
}  // namespace

int pack_topk(ucfp_lane *ctx, const uint64_t *ids, const void *keys32, size_t n, void *records_out) {
    if (n == 0) return UCFP_OK;
    pack_topk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ids, static_cast<const uint32_t *>(keys32), n, static_cast<ulonglong2 *>(records_out));
    count_launch(ctx);
    return check_launch("pack_topk");
}

int merge_packed(ucfp_lane *ctx, const void *records, size_t parts, size_t nq, size_t k, int key_is_f32, int descending, uint64_t *ids_out, void *keys_out) {
    if (key_is_f32) return merge_impl<float>(ctx, static_cast<const uint64_t *>(records), nullptr, parts, nq, k, 1, -INFINITY, ids_out, static_cast<float *>(keys_out));
    return merge_impl<uint32_t>(ctx, static_cast<const uint64_t *>(records), nullptr, parts, nq, k, descending, UINT32_MAX, ids_out, static_cast<uint32_t *>(keys_out));
}

// ---- scan diagnostics ------------------------------------------------------------------------
namespace {
__global__ void add_flags_kernel(const uint32_t *flags, uint32_t nq, unsigned long long *stats) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq && flags[i]) atomicAdd(&stats[0], 1ULL);
}
}  // namespace

int stats_reset(ucfp_lane *ctx) {
    UCFP_TRY(ctx->stats.reserve(64));
    UCFP_CUDA_TRY(cudaMemsetAsync(ctx->stats.ptr, 0, 64, ctx->stream));
    return UCFP_OK;
}

int stats_add_flags(ucfp_lane *ctx, const uint32_t *flags_dev, uint32_t nq) {
    add_flags_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(flags_dev, nq, ctx->stats.as<unsigned long long>());
    count_launch(ctx);
    return check_launch("add_flags");
}

// ---- synthetic data (bench/test support) ---------------------------------------------------
namespace {
__global__ void synth_fill_kernel(uint64_t *dst, uint64_t n, uint64_t seed, uint64_t start) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = splitmix64(seed, start + i);
}
}  // namespace

int synth_fill_u64(ucfp_lane *ctx, uint64_t *dst_dev, uint64_t nwords, uint64_t seed, uint64_t start_word) {
    if (nwords == 0) return UCFP_OK;
    uint64_t blocks = (nwords + 255) / 256;
    uint64_t maxb = (uint64_t)ctx->sm_count * 16;
    if (blocks > maxb) blocks = maxb;
    synth_fill_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(dst_dev, nwords, seed, start_word);
    count_launch(ctx);
    return check_launch("synth_fill");
}

}  // namespace ucfp
