// multihash.cu -- block-hash-aware compare and re-rank over multi bundles (docs/HASH_SPEC.md section 10).
//
// Reference anchors (paths relative to the reference tree): the compare-time MultiHashConfig threaded through
// fingerprint_multi_with (src/modality/image.rs:21-24, 90-104), its fields and defaults (src/server/dto.rs:462-480,
// web/src/lib/docs/api-reference-image.md:51-62).  The 16 block hashes per algorithm that the hashing kernels compute
// and the bundle stores are used here and nowhere else.
//
// A MULTIHASH corpus row is the 51-word hash part of a bundle (ahash[17] | phash[17] | dhash[17]).  Beside the rows the
// corpus keeps a HAMMING64 side corpus of the PHash global hashes (word 17) in the same row order (it borrows the bundle corpus's
// id column), so the coarse pass of a re-rank scan is the ordinary Hamming scan (tensor-core path for batches >= 64), ranking by
// (distance, record id) and reporting ROWS.
#include <math.h>

#include "api_util.cuh"

namespace ucfp {
namespace {

constexpr int kWords = 51;
constexpr int kPhashGlobal = 17;

struct MhCfg { float wa, wp, wd, wg, wb; uint32_t thr; };

__global__ void gather_word_kernel(const uint64_t *__restrict__ rows, uint64_t first, uint64_t n, uint64_t *codes) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) codes[first + i] = rows[(first + i) * kWords + kPhashGlobal];
}

// spec section 10, every operation separately rounded in the order written
__device__ __forceinline__ float blend(const int dg[3], const int m[3], const MhCfg &c) {
    float s[3];
    const float den_gb = __fadd_rn(c.wg, c.wb);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float g = __fdiv_rn((float)(64 - dg[a]), 64.0f), b = __fdiv_rn((float)m[a], 16.0f);
        const float t3 = __fadd_rn(__fmul_rn(c.wg, g), __fmul_rn(c.wb, b));
        s[a] = den_gb == 0.0f ? 0.0f : __fdiv_rn(t3, den_gb);
    }
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(c.wa, s[0]), __fmul_rn(c.wp, s[1])), __fmul_rn(c.wd, s[2]));
    const float den = __fadd_rn(__fadd_rn(c.wa, c.wp), c.wd);
    return den == 0.0f ? 0.0f : __fdiv_rn(num, den);
}

// One warp per pair: lane j < 51 owns word j of both bundles.  Lanes 0 / 17 / 34 hold the global hashes, the other
// 48 lanes the block hashes; a ballot turns "block within the threshold" into three 16-bit popcounts.
__device__ __forceinline__ float warp_score(const uint64_t *x, const uint64_t *y, const MhCfg &c, int lane) {
    uint64_t a0 = 0, b0 = 0, a1 = 0, b1 = 0;
    a0 = x[lane]; b0 = y[lane];                         // words 0..31
    if (lane + 32 < kWords) { a1 = x[lane + 32]; b1 = y[lane + 32]; }   // words 32..50
    const int d0 = __popcll(a0 ^ b0), d1 = __popcll(a1 ^ b1);
    const uint32_t near0 = __ballot_sync(0xFFFFFFFFu, (uint32_t)d0 <= c.thr), near1 = __ballot_sync(0xFFFFFFFFu, (uint32_t)d1 <= c.thr);
    int dg[3], m[3];
    dg[0] = __shfl_sync(0xFFFFFFFFu, d0, 0);
    dg[1] = __shfl_sync(0xFFFFFFFFu, d0, 17);
    dg[2] = __shfl_sync(0xFFFFFFFFu, d1, 34 - 32);
    // block words: ahash 1..16, phash 18..33 (18..31 in the first ballot, 32..33 in the second), dhash 35..50 (3..18 of the second)
    m[0] = __popc(near0 & 0x0001FFFEu);
    m[1] = __popc(near0 & 0xFFFC0000u) + __popc(near1 & 0x00000003u);
    m[2] = __popc(near1 & 0x0007FFF8u);
    return blend(dg, m, c);
}

// candidates of a re-rank: cand_rows[q][j] = row (UINT64_MAX: none) -> packed record {id, score bits}
__global__ void __launch_bounds__(256) rerank_score_kernel(const uint64_t *__restrict__ rows, const uint64_t *__restrict__ ids, uint64_t id_base,
                                                            const uint64_t *__restrict__ queries, const uint64_t *__restrict__ cand_rows, uint32_t nq,
                                                            uint32_t kp, MhCfg cfg, ulonglong2 *out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nq * kp) return;
    const uint32_t q = warp / kp;
    const uint64_t r = cand_rows[warp];
    if (r == UINT64_MAX) { if (lane == 0) out[warp] = make_ulonglong2(UINT64_MAX, (unsigned long long)__float_as_uint(-INFINITY)); return; }
    const float s = warp_score(rows + r * kWords, queries + (uint64_t)q * kWords, cfg, lane);
    if (lane == 0) out[warp] = make_ulonglong2(ids ? ids[r] : id_base + r, (unsigned long long)__float_as_uint(s));
}

__global__ void __launch_bounds__(256) pair_score_kernel(const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, uint64_t n, MhCfg cfg, float *out) {
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const float s = warp_score(a + warp * kWords, b + warp * kWords, cfg, lane);
    if (lane == 0) out[warp] = s;
}

MhCfg make_cfg(const ucfp_multihash_config *c) {
    if (!c) return MhCfg{0.1f, 0.4f, 0.3f, 0.1f, 0.1f, 12u};   // web/src/lib/docs/api-reference-image.md:51-62
    return MhCfg{c->ahash_weight, c->phash_weight, c->dhash_weight, c->global_weight, c->block_weight, c->block_distance_threshold};
}

int check_cfg(const ucfp_multihash_config *c) {
    if (!c) return UCFP_OK;
    const float w[5] = {c->ahash_weight, c->phash_weight, c->dhash_weight, c->global_weight, c->block_weight};
    for (float v : w) UCFP_REQUIRE(v >= 0.0f && v <= 1.0f, UCFP_E_INVALID, "multi-hash weights must lie in [0, 1] (src/server/dto.rs:465-474)");
    UCFP_REQUIRE(c->block_distance_threshold <= 64, UCFP_E_INVALID, "block_distance_threshold must be <= 64");
    return UCFP_OK;
}

}  // namespace

// rows [first, first + n) were written: mirror their PHash global hashes into the coarse corpus and derive its side arrays
int multihash_on_append(ucfp_lane *ctx, ucfp_corpus *c, uint64_t first, uint64_t n) {
    if (n == 0 || !c->coarse) return UCFP_OK;
    gather_word_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(static_cast<const uint64_t *>(c->rows), first, n,
                                                                             static_cast<uint64_t *>(c->coarse->rows));
    count_launch(ctx);
    UCFP_TRY(check_launch("gather_word"));
    c->coarse->size = c->size;   // rows that exist before this call; hamming_on_append extends to first + n itself
    return hamming_on_append(ctx, c->coarse, first, n);
}

}  // namespace ucfp

using namespace ucfp;

extern "C" {

int ucfp_scan_multihash(ucfp_corpus *c, const ucfp_image_hashes *queries, size_t nq, size_t k_prime, size_t k, const ucfp_multihash_config *cfg,
                        uint64_t *ids_out, float *score_out) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    std::shared_lock<std::shared_mutex> rl(c->rw);
    UCFP_LEASE(c->ctx);
    UCFP_REQUIRE(c->kind == UCFP_KIND_MULTIHASH, UCFP_E_STATE, "corpus kind %d cannot serve a multi-hash re-rank", c->kind);
    if (nq == 0 || k == 0) return UCFP_OK;
    UCFP_REQUIRE(queries && ids_out && score_out, UCFP_E_INVALID, "NULL query or output buffer");
    UCFP_REQUIRE(k_prime >= k && k_prime <= 2048, UCFP_E_INVALID, "need k <= k_prime <= 2048 (got k %zu, k_prime %zu)", k, k_prime);
    UCFP_TRY(check_cfg(cfg));
    const MhCfg mc = make_cfg(cfg);
    cudaStream_t st = lane->stream;
    const void *q_dev = nullptr;
    UCFP_TRY(stage_in(lane, lane->q_dev, queries, sizeof(ucfp_image_hashes) * nq, &q_dev));
    // scratch: query codes, coarse candidates (rows + distances), scored records, merged lists
    const size_t n_cand = nq * k_prime;
    UCFP_TRY(lane->mh_a.reserve(8 * nq + 12 * n_cand));
    UCFP_TRY(lane->mh_b.reserve(16 * n_cand + 12 * n_cand));
    uint64_t *q_codes = lane->mh_a.as<uint64_t>(), *cand_rows = q_codes + nq;
    uint32_t *cand_dist = reinterpret_cast<uint32_t *>(cand_rows + n_cand);
    ulonglong2 *scored = lane->mh_b.as<ulonglong2>();
    uint64_t *merged_ids = reinterpret_cast<uint64_t *>(scored + n_cand);
    float *merged_scores = reinterpret_cast<float *>(merged_ids + n_cand);
    UCFP_CUDA_TRY(cudaMemcpy2DAsync(q_codes, 8, static_cast<const uint64_t *>(q_dev) + kPhashGlobal, sizeof(ucfp_image_hashes), 8, nq,
                                    cudaMemcpyDeviceToDevice, st));
    UCFP_TRY(stats_reset(lane));
    // the coarse pass ranks by (distance, RECORD id) -- the candidate set must not depend on where delete / upsert moved a row --
    // but reports rows: it borrows the bundle corpus's id column
    c->coarse->size = c->size; c->coarse->id_mode = c->id_mode; c->coarse->id_base = c->id_base; c->coarse->ids = c->ids;
    UCFP_TRY(hamming_scan(lane, c->coarse, q_codes, nq, k_prime, cand_rows, cand_dist, /*emit_rows=*/true));
    rerank_score_kernel<<<(unsigned)((n_cand * 32 + 255) / 256), 256, 0, st>>>(static_cast<const uint64_t *>(c->rows), c->id_mode == 1 ? c->ids : nullptr, c->id_base,
                                                                              static_cast<const uint64_t *>(q_dev), cand_rows, (uint32_t)nq, (uint32_t)k_prime, mc, scored);
    count_launch(lane);
    UCFP_TRY(check_launch("rerank_score"));
    UCFP_TRY(merge_packed(lane, scored, 1, nq, k_prime, 1, 1, merged_ids, merged_scores));
    // the first k of every sorted k'-list, straight into the caller's buffers (host or device)
    const bool host_out = classify(ids_out) != Mem::Device;
    UCFP_REQUIRE((classify(score_out) != Mem::Device) == host_out, UCFP_E_INVALID, "ids_out and score_out must live in the same memory");
    const cudaMemcpyKind kind = host_out ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    UCFP_CUDA_TRY(cudaMemcpy2DAsync(ids_out, 8 * k, merged_ids, 8 * k_prime, 8 * k, nq, kind, st));
    UCFP_CUDA_TRY(cudaMemcpy2DAsync(score_out, 4 * k, merged_scores, 4 * k_prime, 4 * k, nq, kind, st));
    {
        std::lock_guard<std::mutex> lk(c->ctx->mu);
        for (int i = 0; i < c->ctx->n_lanes; ++i) if (c->ctx->lanes[i] == lane) c->ctx->last_scan_lane = i;
    }
    return finish_call(lane, host_out);
    UCFP_API_END
}

int ucfp_multihash_compare(ucfp_ctx *ctx, const ucfp_image_hashes *a, const ucfp_image_hashes *b, size_t n, const ucfp_multihash_config *cfg,
                           float *score_out) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    UCFP_LEASE(ctx);
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(a && b && score_out, UCFP_E_INVALID, "NULL buffer");
    UCFP_TRY(check_cfg(cfg));
    const void *a_dev = nullptr, *b_dev = nullptr;
    UCFP_TRY(stage_in(lane, lane->mh_a, a, sizeof(ucfp_image_hashes) * n, &a_dev));
    UCFP_TRY(stage_in(lane, lane->mh_b, b, sizeof(ucfp_image_hashes) * n, &b_dev));
    void *out_dev = nullptr;
    bool out_host = false;
    UCFP_TRY(stage_out(lane->out_keys_dev, score_out, 4 * n, &out_dev, &out_host));
    pair_score_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, lane->stream>>>(static_cast<const uint64_t *>(a_dev), static_cast<const uint64_t *>(b_dev), n,
                                                                                make_cfg(cfg), static_cast<float *>(out_dev));
    count_launch(lane);
    UCFP_TRY(check_launch("pair_score"));
    if (out_host) UCFP_TRY(copy_back(lane, score_out, out_dev, 4 * n));
    return finish_call(lane, out_host);
    UCFP_API_END
}

}  // extern "C"
