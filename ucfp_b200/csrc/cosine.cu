// cosine.cu -- brute-force cosine top-k over an HBM-resident embedding corpus (sm_100a).
// Replaces EmbeddedBackend::knn (src/index/embedded/mod.rs:268-360) and its helpers dot_product (:454-472),
// l2_norm (:475), insert_topk (:484-495) of the reference.
//
// The reference scores every row in f32:  score = dot(q, v) / (|q| * |v|), dot = 8 independent f32 lane sums
// over chunks of 8, lanes added left to right, then the scalar remainder.  This file returns exactly those
// f32 scores (compiled with -fmad=false, IEEE sqrt/div) and the exact top-k under (score desc, id asc), but
// it does not compute 2*N*d flops per query on CUDA cores.  Per chunk of the corpus:
//   1. coarse pass on the tensor cores: rows and queries are kept as unit-norm bf16 copies; a tcgen05.mma
//      GEMM (TMA -> 128B-swizzled shared memory -> UMMA, f32 accumulators in TMEM) gives s~(q, r) with
//      |s~ - score| <= kCoarseEps for every pair (bf16 rounding of unit vectors: 2 * 2^-9 plus f32
//      accumulation noise).  The epilogue reads TMEM with tcgen05.ld and appends row r to
//      query q's candidate list iff s~ >= thr_q - kCoarseEps, where thr_q is the exact score of q's current
//      k-th result.  No row that could enter the top-k is ever dropped.
//   2. rescoring: candidates (a few dozen per query) are scored exactly from the f32 rows in the reference's
//      summation order, merged with the kept list, sorted by (score desc, id asc); thr_q tightens.
// Overflowing candidate lists fall back to the cooperative exact selection of topk_select.cuh.
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace ucfp {
namespace {

#include "topk_select.cuh"
#include "sm100_ptx.cuh"

constexpr float kCoarseEps = 0.0078125f;   // 2^-7, see header
constexpr uint32_t kSeedRows = 512;        // kSmallList (topk_select.cuh): list entries the small rescoring launch sorts
constexpr uint32_t kMaxQueriesPerPass = 1024;
constexpr uint64_t kMaxChunkRows = 1ULL << 24;
constexpr uint32_t kCap = 4096;            // candidate rows per query between two rescoring steps
constexpr uint32_t kMaxK = 1024;

// ---- GEMM tile geometry --------------------------------------------------------------------------
constexpr int kTileRows = 128;             // UMMA M: corpus rows per CTA tile (TMEM lanes)
constexpr int kBlockK = 64;                // bf16 elements per K chunk = 128 bytes = one swizzle atom row
constexpr int kStages = 4;
constexpr int kGemmThreads = 192;          // warp 0: TMA, warp 1: MMA, warps 2-5: epilogue
constexpr int kABytes = kTileRows * kBlockK * 2;   // 16 KiB
constexpr int kMaxNTile = 256;             // UMMA N: queries per accumulator tile (TMEM columns)
constexpr int kBBytesMax = kMaxNTile * kBlockK * 2;  // 32 KiB

// ---- exact reference arithmetic --------------------------------------------------------------------
// 8 lanes cooperate on one dot product: lane j owns accumulator j (embedded/mod.rs:458-466), lane 0 then
// folds the eight sums left to right and adds the remainder (:467-471).  `group` = 8 consecutive lanes.
__device__ __forceinline__ float dot8_group(const float *__restrict__ a, const float *__restrict__ b, uint32_t dim, int j,
                                            unsigned group_mask, int group_base) {
    float acc = 0.0f;
    const uint32_t chunks = dim / 8;
    for (uint32_t c = 0; c < chunks; ++c) acc = acc + a[c * 8 + j] * b[c * 8 + j];
    float sum = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) sum = sum + __shfl_sync(group_mask, acc, group_base + l);
    for (uint32_t i = chunks * 8; i < dim; ++i) sum = sum + a[i] * b[i];
    return sum;  // identical on all 8 lanes
}

// one thread does the whole dot product in the same order (exact-selection fallback)
__device__ __forceinline__ float dot8_thread(const float *__restrict__ a, const float *__restrict__ b, uint32_t dim) {
    float accs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t chunks = dim / 8;
    for (uint32_t c = 0; c < chunks; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) accs[j] = accs[j] + a[c * 8 + j] * b[c * 8 + j];
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum = sum + accs[j];
    for (uint32_t i = chunks * 8; i < dim; ++i) sum = sum + a[i] * b[i];
    return sum;
}

// rows (or queries) -> f32 norm (reference arithmetic) + unit-norm bf16 copy padded to dim_pad
__global__ void cosine_prepare_kernel(const float *__restrict__ rows, uint64_t first, uint64_t n, uint32_t dim, uint32_t dim_pad,
                                      float *norm_out, __nv_bfloat16 *unit_out) {
    const int lane = threadIdx.x & 31, j = lane & 7, gbase = lane & ~7;
    const unsigned gmask = 0xFFu << gbase;
    const uint64_t r = first + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8;
    if (r >= first + n) return;  // whole groups leave together
    const float *v = rows + r * dim;
    const float norm = sqrtf(dot8_group(v, v, dim, j, gmask, gbase));
    if (j == 0) norm_out[r] = norm;
    const float inv = norm > 0.0f ? 1.0f / norm : 0.0f;
    __nv_bfloat16 *u = unit_out + r * dim_pad;
    for (uint32_t d = j; d < dim_pad; d += 8) u[d] = __float2bfloat16_rn(d < dim ? v[d] * inv : 0.0f);
}

// ---- per-query selection state ---------------------------------------------------------------------
struct KeptEntry { float score; uint32_t row; };
struct CosState {
    uint32_t *cand;       // [nq][kCap] candidate rows appended by the coarse pass
    uint32_t *count;      // [nq]
    KeptEntry *kept;      // [nq][k] exact results so far, best first
    uint32_t *kept_n;     // [nq]
    float *thr;           // [nq] exact score of the current k-th result (-inf while fewer than k)
    uint32_t *flags;      // [nq]
    uint32_t *big;        // [nq] list too long for the small rescoring launch
};

__global__ void cosine_init_kernel(CosState S, uint32_t nq) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    S.count[i] = 0; S.kept_n[i] = 0; S.thr[i] = -INFINITY; S.flags[i] = 0;
}

// the first `rows` rows become candidates of every query without a coarse pass
__global__ void cosine_seed_kernel(CosState S, uint32_t rows) {
    uint32_t q = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) S.cand[(size_t)q * kCap + r] = r;
    if (r == 0) S.count[q] = rows;
}

__device__ __forceinline__ bool hit_before(float sa, uint64_t ia, float sb, uint64_t ib) { return sa > sb || (sa == sb && ia < ib); }

// One CTA per query: exact scores of the new candidates (8 lanes each), merge with the kept list, sort by
// (score desc, id asc), keep k, publish thr.  The last call writes the result slots.
__global__ void __launch_bounds__(512)
cosine_rescore_kernel(CosState S, uint32_t k, const float *__restrict__ rows, const float *__restrict__ row_norm, uint32_t dim,
                      const float *__restrict__ queries, const float *__restrict__ q_norm, const uint64_t *__restrict__ ids,
                      uint64_t id_base, int final_pass, uint64_t *ids_out, float *score_out, uint32_t smem_entries, int second_launch) {
    extern __shared__ unsigned char sm_raw[];
    const uint32_t q = blockIdx.x;
    const uint32_t n_raw = S.count[q];
    const uint32_t n_new = min(n_raw, kCap);
    const uint32_t n_old = S.kept_n[q];
    const uint32_t n = n_new + n_old;
    // Two launches per step: the first with shared memory for kSmallList entries (16 KiB: many CTAs per SM, one wave) handles
    // the usual short lists and leaves longer ones, marked in S.big, to the second launch (128 KiB, one CTA per SM), whose
    // other CTAs return at once.
    if (!second_launch) {
        const bool big = n > smem_entries;
        if (threadIdx.x == 0) S.big[q] = big ? 1u : 0u;
        if (big) return;
    } else if (!S.big[q]) return;
    uint32_t P = 1;
    while (P < n) P <<= 1;
    uint64_t *s_id = reinterpret_cast<uint64_t *>(sm_raw);
    float *s_sc = reinterpret_cast<float *>(s_id + P);
    uint32_t *s_row = reinterpret_cast<uint32_t *>(s_sc + P);
    const float qn = q_norm[q];
    const float *qv = queries + (size_t)q * dim;

    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) { s_id[i] = UINT64_MAX; s_sc[i] = -INFINITY; s_row[i] = 0xFFFFFFFFu; }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_old; i += blockDim.x) {
        KeptEntry e = S.kept[(size_t)q * k + i];
        s_sc[i] = e.score; s_row[i] = e.row; s_id[i] = ids ? ids[e.row] : id_base + e.row;
    }
    const int lane = threadIdx.x & 31, j = lane & 7, gbase = lane & ~7;
    const unsigned gmask = 0xFFu << gbase;
    const uint32_t groups = blockDim.x / 8;
    for (uint32_t c = threadIdx.x / 8; c < ((n_new + groups - 1) / groups) * groups; c += groups) {
        if (c < n_new) {  // uniform within a group of 8 lanes
            const uint32_t r = S.cand[(size_t)q * kCap + c];
            const float vn = row_norm[r];
            float score = -INFINITY;
            uint64_t id = UINT64_MAX;
            const float dot = dot8_group(qv, rows + (size_t)r * dim, dim, j, gmask, gbase);
            if (vn != 0.0f && qn != 0.0f) {   // zero-norm rows never match (embedded/mod.rs:328-330)
                score = dot / (qn * vn);       // :331
                id = ids ? ids[r] : id_base + r;
            }
            if (j == 0) { s_sc[n_old + c] = score; s_row[n_old + c] = r; s_id[n_old + c] = id; }
        }
    }
    __syncthreads();
    for (uint32_t size = 2; size <= P; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                uint32_t i = 2 * t - (t & (stride - 1)), l = i + stride;
                bool up = ((i & size) == 0);
                float si = s_sc[i], sl = s_sc[l]; uint64_t ii = s_id[i], il = s_id[l];
                bool swap = up ? hit_before(sl, il, si, ii) : hit_before(si, ii, sl, il);
                if (swap) { s_sc[i] = sl; s_sc[l] = si; s_id[i] = il; s_id[l] = ii; uint32_t t2 = s_row[i]; s_row[i] = s_row[l]; s_row[l] = t2; }
            }
            __syncthreads();
        }
    // valid entries (id != NONE) sort before invalid ones: score -inf and id MAX go last
    uint32_t m = 0;
    {   // count valid among the first min(n, k): every thread computes the same value
        uint32_t lim = min(n, k);
        uint32_t lo = 0, hi = lim;  // first index with id == NONE (valid entries form a prefix)
        while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (s_id[mid] != UINT64_MAX) lo = mid + 1; else hi = mid; }
        m = lo;
    }
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) S.kept[(size_t)q * k + i] = KeptEntry{s_sc[i], s_row[i]};
    if (final_pass) {
        for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
            ids_out[(size_t)q * k + i] = i < m ? s_id[i] : UINT64_MAX;
            score_out[(size_t)q * k + i] = i < m ? s_sc[i] : -INFINITY;
        }
    }
    if (threadIdx.x == 0) {
        S.kept_n[q] = m;
        S.count[q] = 0;
        if (n_raw > kCap) S.flags[q] = 1;
        S.thr[q] = fmaxf(S.thr[q], m >= k ? s_sc[k - 1] : -INFINITY);   // never below a bound the probe established (it is made of k real rows too)
    }
}

// ---- probe: a bound before the first chunk ------------------------------------------------------------------
// The exhaustive seed (512 rows x every query, exact) re-read the same 1 MB of rows once per query: 265 us of a 3.2 ms
// shard scan, L2-bound.  The probe asks the tensor pass instead: a coarse launch over the first kProbeRows rows in which
// every warp names, per query, the best row of its 32 (cosine_coarse_kernel, probe != 0), then ONE CTA per query scores those
// kProbeRows / 32 rows exactly.  The k-th best of them is a valid lower bound of the true k-th best score -- k real rows of
// this corpus reach it -- and the main scan starts from it at row 0 with nothing in its lists, so no row is ever
// inserted twice.  Exactness never depends on which rows the probe picked.
constexpr uint32_t kProbeRows = 2048;        // a multiple of the 128-row tile; 64 candidates per query
constexpr uint32_t kProbeRowsMax = 8192;
constexpr uint64_t kProbeMinCorpus = 1ULL << 15;

__global__ void __launch_bounds__(256)
cosine_probe_bound_kernel(CosState S, uint32_t k, uint32_t n_cand, const float *__restrict__ rows, const float *__restrict__ row_norm, uint32_t dim,
                          const float *__restrict__ queries, const float *__restrict__ q_norm) {
    __shared__ float s_sc[kProbeRowsMax / 32];
    const uint32_t q = blockIdx.x;
    const float qn = q_norm[q];
    const float *qv = queries + (size_t)q * dim;
    const int lane = threadIdx.x & 31, j = lane & 7, gbase = lane & ~7;
    const unsigned gmask = 0xFFu << gbase;
    const uint32_t groups = blockDim.x / 8;
    for (uint32_t c = threadIdx.x / 8; c < ((n_cand + groups - 1) / groups) * groups; c += groups) {
        if (c < n_cand) {   // uniform within a group of 8 lanes
            const uint32_t r = S.cand[(size_t)q * kCap + c];
            const float vn = row_norm[r];
            const float dot = dot8_group(qv, rows + (size_t)r * dim, dim, j, gmask, gbase);
            if (j == 0) s_sc[c] = (vn != 0.0f && qn != 0.0f) ? dot / (qn * vn) : -INFINITY;   // the rescoring step's arithmetic
        }
    }
    __syncthreads();
    if (threadIdx.x < n_cand && k <= n_cand) {   // rank by counting: the candidate with exactly k - 1 better ones is the k-th best
        const float mine = s_sc[threadIdx.x];
        uint32_t better = 0;
        for (uint32_t i = 0; i < n_cand; ++i) { const float o = s_sc[i]; better += (o > mine) || (o == mine && i < threadIdx.x); }
        if (better == k - 1 && mine > -INFINITY) S.thr[q] = mine;
    }
}

// ---- the coarse pass: bf16 GEMM tile [128 corpus rows] x [all queries], fused candidate filter ------------
// grid = corpus tiles of the chunk.  Per CTA: for every query tile (n_tile columns) run the K loop through a
// 4-stage TMA ring, accumulate in one of two TMEM stages, and let the epilogue warps compare each score with
// the query's admission bound while the next tile's MMAs are already running.
__global__ void __launch_bounds__(kGemmThreads, 1)
cosine_coarse_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_q,
                     uint64_t row0, uint64_t row_end, uint32_t nq, uint32_t n_tile, uint32_t k_chunks, uint32_t q_groups, CosState S, int probe) {
    extern __shared__ unsigned char smem_raw[];
    // 128B-swizzled operand tiles must start on a 1024-byte boundary of the shared window
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sA = smem;                                   // [kStages][16 KiB]
    unsigned char *sB = smem + kStages * kABytes;               // [kStages][32 KiB]
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + kStages * kBBytesMax);
    uint64_t *empty = full + kStages;
    uint64_t *tfull = empty + kStages;     // [2] accumulator stage ready for the epilogue
    uint64_t *tempty = tfull + 2;          // [2] accumulator stage drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    float *s_tau = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));   // [nq padded to 32] thr - eps

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q_tiles = (nq + n_tile - 1) / n_tile;
    const uint32_t b_bytes = n_tile * kBlockK * 2;
    // persistent: this CTA takes corpus tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the three roles walk the same
    // (tile, query tile, K chunk) sequence, so the TMA ring and the two TMEM stages stay full across tile boundaries
    const uint32_t n_tiles = (uint32_t)((row_end - row0 + kTileRows - 1) / kTileRows);
    // Small chunks (the first of a batch) have fewer row tiles than the GPU has SMs: the query tiles are then split over
    // q_groups CTAs per row tile (grid = tiles x groups).  Large chunks use q_groups == 1 and a persistent grid.
    const uint32_t grp = blockIdx.x % q_groups;
    const uint32_t tile_first = blockIdx.x / q_groups, tile_step = gridDim.x / q_groups;
    const uint32_t qt0 = q_tiles * grp / q_groups, qt1 = q_tiles * (grp + 1) / q_groups;

    for (uint32_t i = threadIdx.x; i < ((nq + 31) & ~31u); i += blockDim.x) s_tau[i] = i < nq ? S.thr[i] - kCoarseEps : INFINITY;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_rows) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_512(tmem_slot);  // one warp allocates all 512 TMEM columns (2 accumulator stages x 256)
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t tile = tile_first; tile < n_tiles; tile += tile_step) {
                const uint64_t tile_row = row0 + (uint64_t)tile * kTileRows;
                for (uint32_t qt = qt0; qt < qt1; ++qt)
                    for (uint32_t kc = 0; kc < k_chunks; ++kc, ++it) {
                        const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], kABytes + b_bytes);
                        tma_load_2d(sA + s * kABytes, &map_rows, &full[s], (int)(kc * kBlockK), (int)tile_row);
                        tma_load_2d(sB + s * kBBytesMax, &map_q, &full[s], (int)(kc * kBlockK), (int)(qt * n_tile));
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane) =====
        // instruction descriptor: D = f32, A = B = bf16, both K-major, N = n_tile, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((n_tile >> 3) << 17) | ((kTileRows >> 4) << 24);
        uint32_t it = 0, acc_it = 0;
        for (uint32_t tile = tile_first; tile < n_tiles; tile += tile_step)
            for (uint32_t qt = qt0; qt < qt1; ++qt, ++acc_it) {
                const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                mbar_wait(&tempty[as], aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * kMaxNTile;
                for (uint32_t kc = 0; kc < k_chunks; ++kc, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(&full[s], ph);
                    tcgen05_fence_after();
                    if (lane == 0) {
                        const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * kABytes));
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * kBBytesMax));
#pragma unroll
                        for (int kk = 0; kk < kBlockK / 16; ++kk)   // UMMA_K = 16 bf16 = 32 bytes = +2 in the address field
                            umma_bf16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kc | kk) ? 1u : 0u);
                        umma_commit(&empty[s]);                      // frees the smem stage when these MMAs retire
                        if (kc == k_chunks - 1) umma_commit(&tfull[as]);
                    }
                    __syncwarp();
                }
            }
    } else {
        // ===== epilogue: 4 warps, warp w reads TMEM lanes 32*(w%4).. (its hardware quadrant) =====
        const uint32_t quad = warp & 3;
        uint32_t acc_it = 0;
        for (uint32_t tile = tile_first; tile < n_tiles; tile += tile_step) {
            const uint64_t my_row = row0 + (uint64_t)tile * kTileRows + quad * 32 + lane;
            const bool valid = my_row < row_end;
            for (uint32_t qt = qt0; qt < qt1; ++qt, ++acc_it) {
                const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                mbar_wait(&tfull[as], aph);
                tcgen05_fence_after();
                const uint32_t cols = min(n_tile, nq - qt * n_tile);
                for (uint32_t c0 = 0; c0 < cols; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((quad * 32u) << 16) + as * kMaxNTile + c0, v);
                    if (probe) {
                        // probe launch (cosine_scan): no bound exists yet.  Every warp reports, per query, the row with the best
                        // COARSE score among its 32 rows: S.cand[q][group of 32 rows].  One REDUX per column on the score's
                        // order-preserving integer image, the low 5 bits traded for the lane; no atomics, no lists.
                        const uint32_t group = (uint32_t)((my_row - row0) >> 5);
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const uint32_t b = v[c];
                            const uint32_t ordered = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
                            const uint32_t best = __reduce_max_sync(0xffffffffu, valid ? ((ordered & ~31u) | (31u - (uint32_t)lane)) : 0u);
                            if (lane == 0 && c0 + c < cols)
                                S.cand[(size_t)(qt * n_tile + c0 + c) * kCap + group] = (uint32_t)(my_row + (31u - (best & 31u)));
                        }
                        continue;
                    }
                    // admission bounds of these 32 queries (padded with +inf): 8 broadcast LDS.128, then one pass
                    // that only records whether ANY score clears its bound -- survivors are rare
                    const float4 *tp = reinterpret_cast<const float4 *>(s_tau + qt * n_tile + c0);
                    float tau[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { float4 t = tp[j]; tau[4 * j] = t.x; tau[4 * j + 1] = t.y; tau[4 * j + 2] = t.z; tau[4 * j + 3] = t.w; }
                    bool any = false;
#pragma unroll
                    for (int c = 0; c < 32; ++c) any |= __uint_as_float(v[c]) >= tau[c];
                    if (any && valid) {
                        // All the list slots are requested first, then filled: an in-order thread that tests each returned
                        // position before asking for the next one pays the ~2 us round trip of a contended atomic once per
                        // survivor -- with the loose bounds of a batch's first chunks that was most of a small launch
                        // (4 096 rows: 85 us, tensor pipe 3 % active).
                        uint32_t pos[32];
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            pos[c] = kCap;
                            if (__uint_as_float(v[c]) >= tau[c]) pos[c] = atomicAdd(&S.count[qt * n_tile + c0 + c], 1u);
                        }
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (pos[c] < kCap) S.cand[(size_t)(qt * n_tile + c0 + c) * kCap + pos[c]] = (uint32_t)my_row;
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[as]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_512(tmem_base);
    }
}

constexpr size_t kGemmSmem = (size_t)kStages * (kABytes + kBBytesMax) + 16 * 8 + 16 + (kMaxQueriesPerPass + 32) * 4 + 1024;

// ---- exact-selection key for flagged queries --------------------------------------------------------------
struct CosineKey {
    static constexpr int kKeyBits = 32;
    static constexpr uint32_t kInvalidKey = 0xFFFFFFFFu;
    const float *rows; const float *row_norm; const float *queries; const float *q_norm; uint32_t dim;
    const float *qv; float qn;
    __device__ void load_query(uint32_t q) { qv = queries + (size_t)q * dim; qn = q_norm[q]; }
    // descending score -> ascending key: order-preserving map of the f32 bits, inverted
    __device__ static uint32_t score_key(float s) {
        uint32_t b = __float_as_uint(s);
        uint32_t ordered = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
        return ~ordered;
    }
    __device__ static uint32_t report(uint32_t key, uint32_t) {
        uint32_t ordered = ~key;
        uint32_t b = (ordered & 0x80000000u) ? (ordered & 0x7FFFFFFFu) : ~ordered;
        return b;  // raw f32 bits, written through a uint32_t view of score_out
    }
    __device__ uint32_t key(uint64_t r) const {
        const float vn = row_norm[r];
        if (vn == 0.0f || qn == 0.0f) return kInvalidKey;
        float s = dot8_thread(qv, rows + r * dim, dim) / (qn * vn);
        uint32_t kk = score_key(s);
        return kk == kInvalidKey ? kInvalidKey - 1 : kk;
    }
};

__global__ void fill_sentinel_f32_kernel(uint64_t *ids_out, float *score_out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { ids_out[i] = UINT64_MAX; score_out[i] = -INFINITY; }
}
__global__ void fix_sentinel_scores_kernel(const uint64_t *ids_out, float *score_out, const uint32_t *flags, uint32_t k, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // exact_select writes UINT32_MAX into empty slots
    if (i < n && flags[i / k] && ids_out[i] == UINT64_MAX) score_out[i] = -INFINITY;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap *map, void *base, uint64_t rows, uint32_t dim_pad, uint32_t box_rows) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        UCFP_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        UCFP_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, UCFP_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cuuint64_t dims[2] = {dim_pad, rows};
    cuuint64_t strides[1] = {(cuuint64_t)dim_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UCFP_REQUIRE(r == CUDA_SUCCESS, UCFP_E_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
    return UCFP_OK;
}

}  // namespace

int cosine_on_append(ucfp_lane *ctx, ucfp_corpus *c, uint64_t first_row, uint64_t n) {
    if (n == 0) return UCFP_OK;
    uint64_t threads = n * 8;
    cosine_prepare_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(
        static_cast<const float *>(c->rows), first_row, n, c->dim, c->dim_pad, c->cos_inv_norm, static_cast<__nv_bfloat16 *>(c->cos_bf16));
    count_launch(ctx);
    return check_launch("cosine_prepare");
}

int cosine_device_init(ucfp_ctx *ctx) {
    UCFP_CUDA_TRY(cudaFuncSetAttribute(cosine_coarse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(cosine_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8192));
    int occ = 0;
    UCFP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, exact_select_kernel<CosineKey>, 256, 0));
    ctx->cos_exact_occ = occ < 1 ? 1 : (occ > 4 ? 4 : occ);
    return UCFP_OK;
}

int cosine_scan(ucfp_lane *ctx, ucfp_corpus *c, const float *q_dev, size_t nq, size_t k, uint64_t *ids_out_dev, float *score_out_dev) {
    cudaStream_t st = ctx->stream;
    const uint64_t N = c->size;
    const uint32_t dim = c->dim, dim_pad = c->dim_pad;
    UCFP_REQUIRE(k <= kMaxK, UCFP_E_UNSUPPORTED, "cosine scan supports k <= %u (got %zu)", kMaxK, k);
    UCFP_REQUIRE(N < (1ULL << 31), UCFP_E_UNSUPPORTED, "cosine corpus limited to 2^31 rows per GPU");
    if (N == 0) {
        size_t tot = nq * k;
        fill_sentinel_f32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ids_out_dev, score_out_dev, tot);
        count_launch(ctx);
        return check_launch("fill_sentinel");
    }
    const float *rows = static_cast<const float *>(c->rows);
    const float *row_norm = c->cos_inv_norm;  // holds |v| (f32, reference arithmetic)
    const uint64_t *ids = c->id_mode == 1 ? c->ids : nullptr;
    CUtensorMap map_rows;
    UCFP_TRY(make_map(&map_rows, c->cos_bf16, c->capacity + 256, dim_pad, kTileRows));

    for (size_t q0 = 0; q0 < nq; q0 += kMaxQueriesPerPass) {
        const uint32_t nqp = (uint32_t)((nq - q0 < kMaxQueriesPerPass) ? nq - q0 : kMaxQueriesPerPass);
        const uint32_t n_tile = nqp >= kMaxNTile ? kMaxNTile : ((nqp + 15) / 16) * 16;
        const uint32_t nq_pad = ((nqp + n_tile - 1) / n_tile) * n_tile;
        // scratch: unit-norm bf16 queries [nq_pad][dim_pad] + norms, selection state
        size_t off_norm = (size_t)nq_pad * dim_pad * 2;
        UCFP_TRY(ctx->qstate.reserve(off_norm + 4 * (size_t)nq_pad + 256));
        __nv_bfloat16 *q_unit = ctx->qstate.as<__nv_bfloat16>();
        float *q_norm = reinterpret_cast<float *>(ctx->qstate.as<unsigned char>() + off_norm);
        UCFP_TRY(ctx->cand.reserve(4 * (size_t)kCap * nqp));
        size_t misc2 = (size_t)nqp * (8 * k + 4 + 4 + 4 + 4 + 4) + 256;
        UCFP_TRY(ctx->cand_count.reserve(misc2));
        CosState S;
        S.cand = ctx->cand.as<uint32_t>();
        S.kept = ctx->cand_count.as<KeptEntry>();
        S.count = reinterpret_cast<uint32_t *>(S.kept + (size_t)nqp * k);
        S.kept_n = S.count + nqp;
        S.thr = reinterpret_cast<float *>(S.kept_n + nqp);
        S.flags = reinterpret_cast<uint32_t *>(S.thr + nqp);
        S.big = S.flags + nqp;
        const float *qp = q_dev + q0 * dim;
        uint64_t *ids_out = ids_out_dev + q0 * k;
        float *score_out = score_out_dev + q0 * k;

        UCFP_CUDA_TRY(cudaMemsetAsync(q_unit, 0, off_norm, st));
        cosine_prepare_kernel<<<(nqp * 8 + 255) / 256, 256, 0, st>>>(qp, 0, nqp, dim, dim_pad, q_norm, q_unit);
        cosine_init_kernel<<<(nqp + 255) / 256, 256, 0, st>>>(S, nqp);
        static const bool env_no_probe = getenv("UCFP_COSINE_NO_PROBE") != nullptr;   // developer switch: the exhaustive 512-row seed
        static const long env_probe_rows = getenv("UCFP_COSINE_PROBE_ROWS") ? atol(getenv("UCFP_COSINE_PROBE_ROWS")) : 0;       // developer knobs
        static const long env_probe_growth = getenv("UCFP_COSINE_PROBE_GROWTH") ? atol(getenv("UCFP_COSINE_PROBE_GROWTH")) : 0;
        const uint32_t probe_rows = env_probe_rows >= 128 && env_probe_rows <= (long)kProbeRowsMax ? (uint32_t)env_probe_rows / 128 * 128 : kProbeRows;
        const uint32_t probe_growth = env_probe_growth > 0 ? (uint32_t)env_probe_growth : 8;
        const bool use_probe = !env_no_probe && N >= kProbeMinCorpus && nqp >= 16 && k <= probe_rows / 64;
        const uint32_t seed = use_probe ? 0u : (uint32_t)(N < kSeedRows ? N : kSeedRows);
        if (!use_probe) cosine_seed_kernel<<<dim3((seed + 255) / 256, nqp), 256, 0, st>>>(S, seed);
        count_launch(ctx, use_probe ? 2 : 3);
        auto rescore = [&](bool final_pass) {
            cosine_rescore_kernel<<<nqp, 256, 16 * kSmallList, st>>>(S, (uint32_t)k, rows, row_norm, dim, qp, q_norm, ids, c->id_base,
                                                                    final_pass ? 1 : 0, ids_out, score_out, kSmallList, 0);
            cosine_rescore_kernel<<<nqp, 512, 16 * 8192, st>>>(S, (uint32_t)k, rows, row_norm, dim, qp, q_norm, ids, c->id_base,
                                                              final_pass ? 1 : 0, ids_out, score_out, 8192, 1);
            count_launch(ctx, 2);
        };
        if (!use_probe) rescore(seed == N);

        CUtensorMap map_q;
        UCFP_TRY(make_map(&map_q, q_unit, nq_pad, dim_pad, n_tile));
        if (use_probe) {
            const uint32_t tiles = probe_rows / kTileRows, q_tiles = (nqp + n_tile - 1) / n_tile;
            uint32_t q_groups = (uint32_t)ctx->sm_count / tiles < q_tiles ? (uint32_t)ctx->sm_count / tiles : q_tiles;
            if (q_groups < 1) q_groups = 1;
            {
                ProfScope ps(ctx, UCFP_PROF_COSINE_SCAN, 2.0 * (double)probe_rows * dim * nqp);
                cosine_coarse_kernel<<<tiles * q_groups, kGemmThreads, kGemmSmem, st>>>(map_rows, map_q, 0, probe_rows, nqp, n_tile, dim_pad / kBlockK, q_groups, S, 1);
            }
            cosine_probe_bound_kernel<<<nqp, 256, 0, st>>>(S, (uint32_t)k, probe_rows / 32, rows, row_norm, dim, qp, q_norm);
            count_launch(ctx, 2);
        }
        uint64_t pos = seed, chunk = use_probe ? (uint64_t)probe_rows * probe_growth : (uint64_t)seed * 8;
        while (pos < N) {
            uint64_t n = (N - pos < chunk) ? N - pos : chunk;
            uint32_t tiles = (uint32_t)((n + kTileRows - 1) / kTileRows);
            {
                ProfScope ps(ctx, UCFP_PROF_COSINE_SCAN, 2.0 * (double)n * dim * nqp);
                uint32_t grid = tiles < (uint32_t)ctx->sm_count ? tiles : (uint32_t)ctx->sm_count;   // 1 CTA per SM (smem), persistent
                uint32_t q_groups = 1;
                const uint32_t q_tiles = (nqp + n_tile - 1) / n_tile;
                if (2 * tiles <= (uint32_t)ctx->sm_count && q_tiles > 1) {   // small chunk: one CTA per (row tile, group of query tiles)
                    q_groups = (uint32_t)ctx->sm_count / tiles < q_tiles ? (uint32_t)ctx->sm_count / tiles : q_tiles;
                    grid = tiles * q_groups;
                }
                cosine_coarse_kernel<<<grid, kGemmThreads, kGemmSmem, st>>>(map_rows, map_q, pos, pos + n, nqp, n_tile, dim_pad / kBlockK, q_groups, S, 0);
            }
            count_launch(ctx);
            pos += n;
            rescore(pos == N);
            chunk = chunk * 8 < kMaxChunkRows ? chunk * 8 : kMaxChunkRows;
        }
        UCFP_TRY(check_launch("cosine scan"));
        CosineKey key{rows, row_norm, qp, q_norm, dim, nullptr, 0.0f};
        UCFP_TRY(stats_add_flags(ctx, S.flags, nqp));
        UCFP_TRY(exact_select_fallback(ctx, c, ctx->owner->cos_exact_occ, key, S.flags, nqp, (uint32_t)k, 0u, ids_out, reinterpret_cast<uint32_t *>(score_out)));
        size_t tot = (size_t)nqp * k;
        fix_sentinel_scores_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ids_out, score_out, S.flags, (uint32_t)k, tot);
        count_launch(ctx);
    }
    return check_launch("cosine scan tail");
}

}  // namespace ucfp
