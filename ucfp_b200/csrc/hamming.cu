// hamming.cu -- HBM-resident brute-force Hamming top-k over 64-bit codes (sm_100a).
//
// Semantics (docs/HASH_SPEC.md section 6; absent from the reference):
//   dist(q, c) = popcount(q ^ c); per query the k smallest under the total order
//   (dist asc, record_id asc).
//
// Structure of one query batch (<= kMaxQueriesPerPass queries per corpus pass):
//   init      per-query state {code, thr = 64, kth_id = MAX}, empty candidate lists
//   seed      the first kSeedRows rows are scored exhaustively into the lists
//   compact   per query: sort list by (dist, id), keep k, publish thr = dist_k, kth_id = id_k
//   scan      geometric chunks of the corpus.  Each thread keeps CPT codes in registers
//             (128-bit coalesced loads, read from HBM exactly once per batch) and walks the
//             query slots staged in shared memory.  Per (query, code) pair the hot loop is
//             2 LOP3 + 1 POPC + 1 min:  f = (c.hi ^ q.hi) | (c.lo ^ q.lo) has
//             popc(f) <= popc(c ^ q), so popc(f) > thr rejects the pair exactly; only
//             survivors compute the true distance and are appended to the query's list
//             when (dist, id) < (thr, kth_id).
//   compact   after every chunk (tightens thr); the last one writes the results.
// A list that overflows its capacity (adversarial duplicates with descending ids) is
// flagged; flagged queries are recomputed by the exact multi-pass selection in
// topk_select.cuh (histogram of distances + radix select on ids), so the result is
// exact for every input.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace ucfp {

namespace {

#include "topk_select.cuh"

constexpr int kScanThreads = 256;
constexpr int kCodesPerThread = 8;                 // 4 x LDG.128 in flight per thread
constexpr int kTileCodes = kScanThreads * kCodesPerThread;
constexpr uint32_t kSeedRows = 2048;               // multiple of 2 (keeps 16-byte alignment of chunk starts)
constexpr uint32_t kMaxQueriesPerPass = 2048;      // 16 B/query of shared memory
constexpr uint64_t kMaxChunkRows = 1ULL << 28;

struct __align__(16) QSlot { uint32_t lo, hi, thr, pad; };

__global__ void hamming_init_kernel(const uint64_t *__restrict__ q, uint32_t nq, QSlot *slots, uint64_t *kth_id,
                                    uint32_t *count, uint32_t *flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    uint64_t c = q[i];
    slots[i] = QSlot{(uint32_t)c, (uint32_t)(c >> 32), 64u, 0u};
    kth_id[i] = UINT64_MAX;
    count[i] = 0;
    flags[i] = 0;
}

// grid (ceil(rows / 256), nq): exhaustive scoring of the first `rows` rows.
__global__ void hamming_seed_kernel(const uint64_t *__restrict__ codes, uint32_t rows, const QSlot *__restrict__ slots,
                                    uint64_t *cand, uint32_t *count, uint32_t cap) {
    uint32_t q = blockIdx.y;
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) {
        QSlot s = slots[q];
        uint64_t c = codes[r];
        uint32_t d = __popc((uint32_t)c ^ s.lo) + __popc((uint32_t)(c >> 32) ^ s.hi);
        cand[(size_t)q * cap + r] = ((uint64_t)d << 40) | r;
    }
    if (r == 0) count[q] = rows;
}

__device__ __forceinline__ uint4 ldg_stream_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// f = (a ^ b) | c in one LOP3 (immLut 0xBE); kept opaque so the compiler does not split it to share
// the XOR with the (cold) exact-distance path.
__device__ __forceinline__ uint32_t xor_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xBE;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Cold path of the scan: exact distances for the 8 codes of one thread against one query; rows with
// (dist, id) < (thr, kth_id) are appended to the query's candidate list.  Out of line so that the hot loop's
// register allocation is not shaped by it, and with the codes passed BY VALUE (in registers): handing the arrays
// over by reference spills them to local memory on every tile, and those stores halve the streaming bandwidth
// (measured: 4.1 vs 7.2 TB/s at one query per pass, scripts/micro/stream_bw2.cu).
__device__ __noinline__ void hamming_survivors(uint4 c01, uint4 c23, uint4 c45, uint4 c67, uint4 s, uint32_t q, uint64_t tile_row,
                                               uint64_t row_end, const uint64_t *__restrict__ ids, uint64_t id_base,
                                               const uint64_t *__restrict__ kth_id, uint64_t *cand, uint32_t *count,
                                               uint32_t cap) {
    const uint32_t lo[kCodesPerThread] = {c01.x, c01.z, c23.x, c23.z, c45.x, c45.z, c67.x, c67.z};
    const uint32_t hi[kCodesPerThread] = {c01.y, c01.w, c23.y, c23.w, c45.y, c45.w, c67.y, c67.w};
    // most calls end here: the one-POPC bound let the pair through but no true distance is within thr
    uint32_t d[kCodesPerThread], dmin = 64;
#pragma unroll
    for (int c = 0; c < kCodesPerThread; ++c) { d[c] = __popc(lo[c] ^ s.x) + __popc(hi[c] ^ s.y); dmin = min(dmin, d[c]); }
    if (dmin > s.z) return;
    const uint64_t kid = kth_id[q];
#pragma unroll
    for (int c = 0; c < kCodesPerThread; ++c) {
        uint64_t r = tile_row + 2ull * ((c >> 1) * kScanThreads + threadIdx.x) + (c & 1);
        if (d[c] <= s.z && r < row_end) {
            uint64_t id = ids ? ids[r] : id_base + r;
            if (d[c] < s.z || id < kid) {
                uint32_t pos = atomicAdd(&count[q], 1u);
                if (pos < cap) cand[(size_t)q * cap + pos] = ((uint64_t)d[c] << 40) | r;
            }
        }
    }
}

// Scans rows [row0, row0 + nrows) of the corpus against the nq staged queries.
__global__ void __launch_bounds__(kScanThreads, 4)
hamming_scan_kernel(const uint64_t *__restrict__ codes, const uint64_t *__restrict__ ids, uint64_t id_base,
                    uint64_t row0, uint64_t nrows, const QSlot *__restrict__ slots, const uint64_t *__restrict__ kth_id,
                    uint32_t nq, uint32_t q_groups, uint64_t *cand, uint32_t *count, uint32_t cap) {
    extern __shared__ uint4 sq[];  // query slots of this CTA's group
    // Small chunks (the first few of a batch, where the admission bounds are still loose and the cold path runs
    // often) have fewer tiles than the GPU has CTA slots: the queries are then split into q_groups groups and the
    // grid is tiles x groups, one (tile, group) pair per CTA.  Large chunks use q_groups == 1 and a persistent grid.
    const uint32_t grp = q_groups > 1 ? blockIdx.x % q_groups : 0;
    const uint32_t q_lo = (uint32_t)((uint64_t)nq * grp / q_groups), q_hi = (uint32_t)((uint64_t)nq * (grp + 1) / q_groups);
    const uint32_t nq_mine = q_hi - q_lo;
    for (uint32_t i = threadIdx.x; i < nq_mine; i += kScanThreads) sq[i] = reinterpret_cast<const uint4 *>(slots)[q_lo + i];
    __syncthreads();

    const uint64_t row_end = row0 + nrows;
    const uint64_t ntiles = (nrows + kTileCodes - 1) / kTileCodes;
    const uint64_t tile0 = q_groups > 1 ? blockIdx.x / q_groups : blockIdx.x;
    const uint64_t tile_step = q_groups > 1 ? ntiles : gridDim.x;
    for (uint64_t tile = tile0; tile < ntiles; tile += tile_step) {
        const uint64_t tile_row = row0 + tile * kTileCodes;  // even by construction
        uint32_t lo[kCodesPerThread], hi[kCodesPerThread];
#pragma unroll
        for (int j = 0; j < kCodesPerThread / 2; ++j) {
            uint64_t r = tile_row + 2ull * (j * kScanThreads + threadIdx.x);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (r + 1 < row_end) v = ldg_stream_v4(reinterpret_cast<const uint4 *>(codes + r));
            else if (r < row_end) { uint64_t c = codes[r]; v.x = (uint32_t)c; v.y = (uint32_t)(c >> 32); }
            lo[2 * j] = v.x; hi[2 * j] = v.y; lo[2 * j + 1] = v.z; hi[2 * j + 1] = v.w;
        }
#pragma unroll 2
        for (uint32_t q = 0; q < nq_mine; ++q) {
            const uint4 s = sq[q];  // broadcast LDS.128: {lo, hi, thr, -}
            uint32_t m = 64;
#pragma unroll
            for (int c = 0; c < kCodesPerThread; ++c) {
                uint32_t f = xor_or(hi[c], s.y, lo[c] ^ s.x);
                m = min(m, (uint32_t)__popc(f));
            }
            if (m <= s.z)  // rare: at least one code of this thread may be within the threshold
                hamming_survivors(make_uint4(lo[0], hi[0], lo[1], hi[1]), make_uint4(lo[2], hi[2], lo[3], hi[3]),
                                  make_uint4(lo[4], hi[4], lo[5], hi[5]), make_uint4(lo[6], hi[6], lo[7], hi[7]), s, q_lo + q,
                                  tile_row, row_end, ids, id_base, kth_id, cand, count, cap);
        }
    }
}

// exact-selection key for flagged queries: the true distance of one row
struct HammingKey {
    static constexpr int kKeyBits = 8;
    static constexpr uint32_t kInvalidKey = 0xFFFFFFFFu;
    __device__ static uint32_t report(uint32_t key, uint32_t flip) { return flip ? flip - key : key; }
    const uint64_t *codes; const QSlot *slots; uint32_t lo, hi;
    __device__ void load_query(uint32_t q) { QSlot s = slots[q]; lo = s.lo; hi = s.hi; }
    __device__ uint32_t key(uint64_t r) const {
        uint64_t c = codes[r];
        return __popc((uint32_t)c ^ lo) + __popc((uint32_t)(c >> 32) ^ hi);
    }
};

}  // namespace

int hamming_scan(ucfp_corpus *c, const uint64_t *q_dev, size_t nq, size_t k, uint64_t *ids_out_dev,
                 uint32_t *dist_out_dev) {
    ucfp_ctx *ctx = c->ctx;
    cudaStream_t st = ctx->stream;
    const uint64_t N = c->size;
    UCFP_REQUIRE(k <= 2048, UCFP_E_UNSUPPORTED, "hamming scan supports k <= 2048 (got %zu)", k);
    UCFP_REQUIRE(N <= kRowMask, UCFP_E_UNSUPPORTED, "corpus too large");
    if (N == 0) {
        size_t tot = nq * k;
        fill_sentinel_u32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ids_out_dev, dist_out_dev, tot);
        count_launch(ctx);
        return check_launch("fill_sentinel");
    }
    uint32_t cap = 4096;
    while (cap < 4 * k) cap <<= 1;
    const uint64_t *codes = static_cast<const uint64_t *>(c->rows);
    const uint64_t *ids = c->id_mode == 1 ? c->ids : nullptr;

    // function attributes are per device: set them on every call (a host-side table write)
    UCFP_CUDA_TRY(cudaFuncSetAttribute(compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8192));
    int scan_occ = 0;
    UCFP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&scan_occ, hamming_scan_kernel, kScanThreads,
                                                                 sizeof(QSlot) * kMaxQueriesPerPass));
    if (scan_occ < 1) scan_occ = 1;

    for (size_t q0 = 0; q0 < nq; q0 += kMaxQueriesPerPass) {
        const uint32_t nqp = (uint32_t)((nq - q0 < kMaxQueriesPerPass) ? nq - q0 : kMaxQueriesPerPass);
        UCFP_TRY(ctx->qstate.reserve(sizeof(QSlot) * nqp + sizeof(uint64_t) * nqp));
        UCFP_TRY(ctx->cand.reserve(sizeof(uint64_t) * (size_t)cap * nqp));
        UCFP_TRY(ctx->cand_count.reserve(sizeof(uint32_t) * nqp));
        UCFP_TRY(ctx->flags.reserve(sizeof(uint32_t) * (nqp + 1)));
        QSlot *slots = ctx->qstate.as<QSlot>();
        uint64_t *kth = reinterpret_cast<uint64_t *>(slots + nqp);
        uint64_t *cand = ctx->cand.as<uint64_t>();
        uint32_t *count = ctx->cand_count.as<uint32_t>();
        uint32_t *flags = ctx->flags.as<uint32_t>();
        uint64_t *ids_out = ids_out_dev + q0 * k;
        uint32_t *dist_out = dist_out_dev + q0 * k;

        hamming_init_kernel<<<(nqp + 255) / 256, 256, 0, st>>>(q_dev + q0, nqp, slots, kth, count, flags);
        const uint32_t seed = (uint32_t)(N < kSeedRows ? N : kSeedRows);
        hamming_seed_kernel<<<dim3((seed + 255) / 256, nqp), 256, 0, st>>>(codes, seed, slots, cand, count, cap);
        count_launch(ctx, 2);

        SelectState sel{cand, count, &slots[0].thr, 4, kth, flags, cap};
        auto compact = [&](bool final_pass) {
            // smem: 16 B per element of the padded list; lists hold <= cap entries
            compact_kernel<<<nqp, 512, 16 * (size_t)cap, st>>>(sel, (uint32_t)k, ids, c->id_base, final_pass ? 1 : 0, 0u,
                                                              ids_out, dist_out);
            count_launch(ctx);
        };
        compact(seed == N);

        // small batches are HBM-bound: fewer, larger chunks; large batches tighten thr more often
        const bool streaming = nqp <= 16;   // HBM-bound regime: fewer and larger chunks
        static const long env_growth = getenv("UCFP_HAMMING_GROWTH") ? atol(getenv("UCFP_HAMMING_GROWTH")) : 0;
        const uint64_t growth = streaming ? 64 : (env_growth > 1 ? (uint64_t)env_growth : 8);
        const uint64_t max_chunk = streaming ? (1ULL << 40) : kMaxChunkRows;
        uint64_t pos = seed, chunk = (uint64_t)seed * growth;
        while (pos < N) {
            uint64_t n = (N - pos < chunk) ? N - pos : chunk;
            uint64_t ntiles = (n + kTileCodes - 1) / kTileCodes;
            const uint64_t slots_full = (uint64_t)ctx->sm_count * scan_occ;
            uint64_t grid = slots_full, q_groups = 1;
            if (ntiles < slots_full) {   // small chunk: split the queries so that every SM gets work
                q_groups = (slots_full + ntiles - 1) / ntiles;
                const uint64_t max_groups = (nqp + 7) / 8;   // at least 8 queries per group
                if (q_groups > max_groups) q_groups = max_groups;
                grid = ntiles * q_groups;
            }
            {
                ProfScope ps(ctx, UCFP_PROF_HAMMING_SCAN, 8.0 * (double)n * nqp);
                hamming_scan_kernel<<<(unsigned)grid, kScanThreads, sizeof(QSlot) * nqp, st>>>(
                    codes, ids, c->id_base, pos, n, slots, kth, nqp, (uint32_t)q_groups, cand, count, cap);
            }
            count_launch(ctx);
            pos += n;
            compact(pos == N);
            chunk = chunk * growth < max_chunk ? chunk * growth : max_chunk;
        }
        UCFP_TRY(check_launch("hamming scan"));
        // exact recomputation of any query whose candidate list overflowed (device-side decision, no host sync)
        UCFP_TRY(stats_add_flags(ctx, flags, nqp));
        UCFP_TRY(exact_select_fallback(c, HammingKey{codes, slots, 0, 0}, flags, nqp, (uint32_t)k, 0u, ids_out, dist_out));
    }
    return UCFP_OK;
}

}  // namespace ucfp
