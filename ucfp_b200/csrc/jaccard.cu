// jaccard.cu -- MinHash-128 signature-equality Jaccard top-k over an HBM-resident corpus (sm_100a).
//
// Semantics (docs/HASH_SPEC.md section 7; absent from the reference):
//   matches(q, r) = #{i < 128 : q.slot[i] == r.slot[i]},  J^ = matches / 128
//   per query the k best under the total order (matches desc, record_id asc).
// Rows are the 1024-byte slot payload of txtfp's MinHashSig<128> (src/modality/text.rs:200-204).
//
// A brute-force pass reads 1 KiB per row.  Instead the corpus keeps, next to the rows, a 128-byte SKETCH per
// row -- the low byte of every slot -- built once at append time.  byte_matches(q, r) >= matches(q, r), so
// the scan streams only the sketches (8x less HBM traffic), counts equal bytes with SWAR arithmetic
// (4 slots per 32-bit word), and only rows whose upper bound can still enter the query's top-k are
// verified against the full 1 KiB row by the whole warp.  Results are exact.
// Selection (thresholds that tighten chunk by chunk, compaction, overflow fallback) is shared with the
// Hamming scan: topk_select.cuh, key = 128 - matches.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace ucfp {
namespace {

#include "topk_select.cuh"

constexpr int kSlots = 128;
constexpr int kSketchWords = 32;                 // 128 slots x 1 byte
constexpr int kScanThreads = 256;
constexpr uint32_t kSeedRows = 1024;
constexpr uint32_t kMaxQueriesPerPass = 256;     // 2 x 128 B of sketch planes + 12 B of bound per query in shared memory
constexpr uint64_t kMaxChunkRows = 1ULL << 26;
constexpr uint32_t kLocalVerifyMax = 8;          // survivors with <= this many agreeing bytes verify lane-locally

// Sketch layout: tiles of 32 rows; word j (slots 4j..4j+3) of row r lives at
// (r / 32) * 1024 + j * 32 + (r % 32), so that a warp reads word j of 32 consecutive rows as one 128-byte line.
__device__ __forceinline__ size_t sketch_index(uint64_t row, int j) { return (row >> 5) * 1024 + (size_t)j * 32 + (row & 31); }
// 32-bit words of one sketch plane as allocated by ucfp_corpus_create (whole 256-row tiles stay readable)
inline size_t sketch_plane_words(uint64_t capacity) { return 32 * ((capacity + 31) / 32 * 32 + 512); }

// plane 0 = byte 0 of every slot (streamed by the scan), plane 1 = byte 1 (consulted only for rows that survive plane 0
// while the bounds are still loose: it turns 255 of 256 chance collisions away without touching the 1 KiB row)
__global__ void sketch_build_kernel(const uint64_t *__restrict__ sigs, uint32_t *sketch, uint32_t *sketch1, uint64_t first, uint64_t n) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * kSketchWords) return;
    uint64_t row = first + idx / kSketchWords;
    int j = (int)(idx % kSketchWords);
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(sigs + row * kSlots + 4 * j);
    ulonglong2 a = p[0], b = p[1];
    uint32_t w = (uint32_t)(a.x & 255) | (uint32_t)(a.y & 255) << 8 | (uint32_t)(b.x & 255) << 16 | (uint32_t)(b.y & 255) << 24;
    sketch[sketch_index(row, j)] = w;
    uint32_t w1 = (uint32_t)((a.x >> 8) & 255) | (uint32_t)((a.y >> 8) & 255) << 8 | (uint32_t)((b.x >> 8) & 255) << 16 | (uint32_t)((b.y >> 8) & 255) << 24;
    sketch1[sketch_index(row, j)] = w1;
}

// query sketches, row-major [q][32], plus per-query selection state
__global__ void jaccard_init_kernel(const uint64_t *__restrict__ q, uint32_t nq, uint32_t *qsketch, uint32_t *thr, uint64_t *kth_id,
                                    uint32_t *count, uint32_t *flags) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nq * kSketchWords) return;
    uint32_t qi = idx / kSketchWords, j = idx % kSketchWords;
    const uint64_t *p = q + (size_t)qi * kSlots + 4 * j;
    qsketch[idx] = (uint32_t)(p[0] & 255) | (uint32_t)(p[1] & 255) << 8 | (uint32_t)(p[2] & 255) << 16 | (uint32_t)(p[3] & 255) << 24;
    qsketch[(size_t)nq * kSketchWords + idx] = (uint32_t)((p[0] >> 8) & 255) | (uint32_t)((p[1] >> 8) & 255) << 8 |
                                               (uint32_t)((p[2] >> 8) & 255) << 16 | (uint32_t)((p[3] >> 8) & 255) << 24;   // plane 1
    if (j == 0) { thr[qi] = 128; kth_id[qi] = UINT64_MAX; count[qi] = 0; flags[qi] = 0; }
}

// exact matches of one (query, row) pair, computed by a full warp: lane l compares slots 4l..4l+3
__device__ __forceinline__ uint32_t warp_matches(const uint64_t *__restrict__ row, const uint64_t *__restrict__ qsig, int lane) {
    const ulonglong2 *rp = reinterpret_cast<const ulonglong2 *>(row + 4 * lane);
    const ulonglong2 *qp = reinterpret_cast<const ulonglong2 *>(qsig + 4 * lane);
    ulonglong2 r0 = __ldg(rp), r1 = __ldg(rp + 1), q0 = __ldg(qp), q1 = __ldg(qp + 1);
    uint32_t m = (r0.x == q0.x) + (r0.y == q0.y) + (r1.x == q1.x) + (r1.y == q1.y);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m += __shfl_xor_sync(0xffffffffu, m, s);
    return m;
}

// grid (ceil(rows / 8), nq), 256 threads: every warp scores one of the first `rows` rows exhaustively.
__global__ void jaccard_seed_kernel(const uint64_t *__restrict__ sigs, uint32_t rows, const uint64_t *__restrict__ q,
                                    uint64_t *cand, uint32_t *count, uint32_t cap) {
    const uint32_t qi = blockIdx.y, lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r < rows) {
        uint32_t m = warp_matches(sigs + (size_t)r * kSlots, q + (size_t)qi * kSlots, lane);
        if (lane == 0) cand[(size_t)qi * cap + r] = ((uint64_t)(128u - m) << 40) | r;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) count[qi] = rows;
}

// SWAR byte compare: 0x80 in every byte where a and b DIFFER (zero-byte detection on a ^ b, inverted)
__device__ __forceinline__ uint32_t ne_bytes_flags(uint32_t a, uint32_t b) {
    uint32_t x = a ^ b;
    uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return (t | x) & 0x80808080u;
}
// Differing bytes of 8 word pairs as ONE bitmap (one popcount counts them): word j's flags sit at bit 7 of each byte;
// g = (g >> 1) + flags moves the older flags down to bits 6..0, so after 8 words every flag has its own bit.
__device__ __forceinline__ uint32_t ne_bytes_8words(const uint32_t *w, uint4 q0, uint4 q1) {
    uint32_t g = ne_bytes_flags(w[0], q0.x);
    g = (g >> 1) + ne_bytes_flags(w[1], q0.y);
    g = (g >> 1) + ne_bytes_flags(w[2], q0.z);
    g = (g >> 1) + ne_bytes_flags(w[3], q0.w);
    g = (g >> 1) + ne_bytes_flags(w[4], q1.x);
    g = (g >> 1) + ne_bytes_flags(w[5], q1.y);
    g = (g >> 1) + ne_bytes_flags(w[6], q1.z);
    g = (g >> 1) + ne_bytes_flags(w[7], q1.w);
    return g;   // bit 8 b + j set <=> byte b of word j differs
}

// kRescan: the launch serves the FLAGGED queries only (a re-scan round after a candidate list overflowed, topk_select.cuh); with no
// flag set every CTA returns at once.
template <bool kRescan>
__global__ void __launch_bounds__(kScanThreads)
jaccard_scan_kernel(const uint32_t *__restrict__ sketch, const uint32_t *__restrict__ sketch1, const uint64_t *__restrict__ sigs,
                    const uint64_t *__restrict__ ids, uint64_t id_base, uint64_t row0, uint64_t nrows, const uint64_t *__restrict__ q_all,
                    const uint32_t *__restrict__ qsketch, uint32_t nq_all, uint32_t q_groups, SelectState S) {
    extern __shared__ uint32_t smem[];
    // Small chunks (the first of a batch) have fewer row tiles than the GPU has CTA slots: the queries are then split into
    // q_groups groups and the grid is tiles x groups.  Large chunks use q_groups == 1 and a persistent grid.
    const uint32_t grp = q_groups > 1 ? blockIdx.x % q_groups : 0;
    const uint32_t q_lo = (uint32_t)((uint64_t)nq_all * grp / q_groups), q_hi = (uint32_t)((uint64_t)nq_all * (grp + 1) / q_groups);
    uint32_t nq = q_hi - q_lo;
    uint32_t *sidx = smem;                      // kRescan: [nq_all] indices of the flagged queries (the arrays below follow it)
    if constexpr (kRescan) {
        __shared__ uint32_t s_listed;
        if (threadIdx.x < 32) {
            const uint32_t listed = list_flagged_queries(S.flags, nq_all, [&](uint32_t pos, uint32_t q) { sidx[pos] = q; });
            if (threadIdx.x == 0) s_listed = listed;
        }
        __syncthreads();
        nq = s_listed;
        if (nq == 0) return;
    }
    uint32_t *sq = smem + (kRescan ? (nq_all + 3u) & ~3u : 0u);   // [nq][32] query sketches, plane 0 (byte 0 of every slot); read as uint4: 16-byte aligned
    uint32_t *sq1 = sq + (size_t)nq * kSketchWords;     // [nq][32] plane 1 (byte 1)
    uint32_t *sthr = sq1 + (size_t)nq * kSketchWords;   // [nq] admission bound (key = 128 - matches)
    uint64_t *skid = reinterpret_cast<uint64_t *>(smem + (((sthr - smem) + nq + 1) & ~(size_t)1));  // [nq] id of the current k-th result
    auto query_of = [&](uint32_t qi) { return kRescan ? sidx[qi] : q_lo + qi; };
    for (uint32_t i = threadIdx.x; i < nq * kSketchWords; i += kScanThreads) {
        const uint32_t qr = query_of(i / kSketchWords), j = i % kSketchWords;
        sq[i] = qsketch[(size_t)qr * kSketchWords + j];
        sq1[i] = qsketch[(size_t)(nq_all + qr) * kSketchWords + j];
    }
    for (uint32_t i = threadIdx.x; i < nq; i += kScanThreads) { sthr[i] = S.thr[(size_t)query_of(i) * S.thr_stride]; skid[i] = S.kth_id[query_of(i)]; }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const uint64_t row_end = row0 + nrows;
    const uint64_t ntiles = (nrows + kScanThreads - 1) / kScanThreads;   // row0 is a multiple of 32
    const uint64_t tile0 = q_groups > 1 ? blockIdx.x / q_groups : blockIdx.x;
    const uint64_t tile_step = q_groups > 1 ? ntiles : gridDim.x;
    for (uint64_t tile = tile0; tile < ntiles; tile += tile_step) {
        const uint64_t row = row0 + tile * kScanThreads + threadIdx.x;
        const bool valid = row < row_end;
        uint32_t w[kSketchWords];
        const uint32_t *sp = sketch + sketch_index(row & ~31ULL, 0) + lane;  // padded allocation: always readable
#pragma unroll
        for (int j = 0; j < kSketchWords; ++j) w[j] = __ldg(sp + j * 32);
        const uint64_t id = valid ? (ids ? ids[row] : id_base + row) : UINT64_MAX;

        for (uint32_t qi = 0; qi < nq; ++qi) {
            const uint4 *qv = reinterpret_cast<const uint4 *>(sq + (size_t)qi * kSketchWords);
            const uint32_t thr = sthr[qi];
            const uint32_t need = 128u - thr;               // byte matches a row needs to stay in the race
            uint32_t bm = 0;                                // byte matches so far: an upper bound on slot matches
            uint32_t ne[4];                                 // per group: bitmap of differing bytes (bit 8 b + j = byte b of word j)
            bool alive = true;
            // four groups of 8 words (32 slots).  After each of the first three, the warp stops as soon as no lane
            // can still reach `need` even if every remaining byte matched -- exact branch-and-bound, and with a
            // meaningful bound almost every (row, query) pair ends after the first group.
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                ne[g] = ne_bytes_8words(&w[8 * g], qv[2 * g], qv[2 * g + 1]);      // two broadcast LDS.128
                bm += 32u - (uint32_t)__popc(ne[g]);
                if (g < 3 && !__any_sync(0xffffffffu, valid && bm + (96u - 32u * g) >= need)) { alive = false; break; }
            }
            if (!alive) continue;
            const uint32_t lower = 128u - bm;               // lower bound on the key
            const uint64_t kid = skid[qi];
            // (key, id) can beat the current k-th only if (lower, id) < (thr, kth_id)
            const bool hit = valid && (lower < thr || (lower == thr && id < kid));
            if (__any_sync(0xffffffffu, hit)) {
                // Survivors with only a few agreeing bytes (chance collisions: 1/256 per slot) are settled by their own
                // lane: only the slots whose low bytes agree can match, so compare just those full 64-bit slots.
                const bool local = hit && bm <= kLocalVerifyMax;
                if (local) {
                    // only the slots whose low bytes agree can match: their positions are the zero bits of ne[].  Byte 1 of the
                    // slot (sketch plane 1, the tile's 4 KiB stay in L1 across the query loop) settles 255 of 256 chance
                    // collisions; the 1 KiB row is read only when both bytes agree.
                    const uint64_t *rp = sigs + row * kSlots;
                    const uint64_t *qp = q_all + (size_t)query_of(qi) * kSlots;
                    uint32_t m = 0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t eq = ~ne[g];
                        while (eq) {
                            const int p = __ffs(eq) - 1;
                            eq &= eq - 1;
                            const int j = 8 * g + (p & 7), b = p >> 3, slot = 4 * j + b;
                            const uint32_t w1 = __ldg(sketch1 + sketch_index(row, j)), q1 = sq1[(size_t)qi * kSketchWords + j];
                            if ((((w1 ^ q1) >> (8 * b)) & 255u) == 0) m += (__ldg(rp + slot) == __ldg(qp + slot));
                        }
                    }
                    const uint32_t key = 128u - m;
                    if (key < thr || (key == thr && id < kid)) {
                        cand_append(S.cand, S.count, S.cap, query_of(qi), ((uint64_t)key << 40) | row);
                    }
                }
                // rows with many agreeing bytes (real neighbours): the whole warp verifies one row at a time
                unsigned hits = __ballot_sync(0xffffffffu, hit && !local);
                while (hits) {
                    const int src = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const uint64_t srow = __shfl_sync(0xffffffffu, row, src);
                    const uint64_t sid = __shfl_sync(0xffffffffu, id, src);
                    const uint32_t m = warp_matches(sigs + srow * kSlots, q_all + (size_t)query_of(qi) * kSlots, lane);
                    const uint32_t key = 128u - m;
                    if (lane == 0 && (key < thr || (key == thr && sid < kid))) {
                        cand_append(S.cand, S.count, S.cap, query_of(qi), ((uint64_t)key << 40) | srow);
                    }
                }
            }
        }
    }
}

// ---- probe: a cheap first look for real neighbours --------------------------------------------------------------------
// Until a query's list holds k good rows its bound is loose, and a chunk scanned under a loose bound costs ~10x more per
// row (no early exit after the first 32 slots, every chance byte collision is verified): on a 6.25 M-row shard that was
// 3.4 of 7.4 ms.  The probe walks the first kProbeRows rows looking at the first 32 slots only and verifies, with the
// whole warp, just the rows in which >= kProbeMinBytes of those 32 low bytes agree (a row sharing a quarter of its slots
// with the query; chance: 32 choose 8 / 256^8).  Its hits go to lists of their own; their k-th best (key, id) is a valid
// upper bound of the true k-th best -- it is made of k real rows of this corpus -- and the main scan starts from it
// (adopt_probe_bound_kernel).  The lists of the main scan never see a probe entry, so nothing is inserted twice; a query
// without k such neighbours keeps its loose bound and behaves as before.
constexpr uint64_t kProbeRows = 1ULL << 19;
constexpr uint64_t kProbeMinCorpus = 1ULL << 17;   // below this the whole scan is cheaper than a probe
constexpr uint32_t kProbeMinBytes = 8;
constexpr int kProbeWords = 8;                     // sketch words of the first 32 slots

__global__ void jaccard_probe_init_kernel(uint32_t nq, uint32_t *pthr, uint64_t *pkid, uint32_t *pcount, uint32_t *pflags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) { pthr[i] = 128; pkid[i] = UINT64_MAX; pcount[i] = 0; pflags[i] = 0; }
}

__global__ void __launch_bounds__(kScanThreads)
jaccard_probe_kernel(const uint32_t *__restrict__ sketch, const uint64_t *__restrict__ sigs, uint64_t row0, uint64_t nrows,
                     const uint64_t *__restrict__ q, const uint32_t *__restrict__ qsketch, uint32_t nq,
                     uint64_t *pcand, uint32_t *pcount, uint32_t cap) {
    extern __shared__ uint32_t smem[];   // [nq][8]: plane-0 sketch words of the first 32 slots
    for (uint32_t i = threadIdx.x; i < nq * kProbeWords; i += kScanThreads) smem[i] = qsketch[(size_t)(i / kProbeWords) * kSketchWords + i % kProbeWords];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t row_end = row0 + nrows;
    const uint64_t ntiles = (nrows + kScanThreads - 1) / kScanThreads;   // row0 is a multiple of 32
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t row = row0 + tile * kScanThreads + threadIdx.x;
        const bool valid = row < row_end;
        uint32_t w[kProbeWords];
        const uint32_t *sp = sketch + sketch_index(row & ~31ULL, 0) + lane;   // padded allocation: always readable
#pragma unroll
        for (int j = 0; j < kProbeWords; ++j) w[j] = __ldg(sp + j * 32);
        for (uint32_t qi = 0; qi < nq; ++qi) {
            const uint4 *qv = reinterpret_cast<const uint4 *>(smem + (size_t)qi * kProbeWords);
            const uint32_t bm = 32u - (uint32_t)__popc(ne_bytes_8words(w, qv[0], qv[1]));
            unsigned hits = __ballot_sync(0xffffffffu, valid && bm >= kProbeMinBytes);
            while (hits) {
                const int src = __ffs(hits) - 1;
                hits &= hits - 1;
                const uint64_t srow = __shfl_sync(0xffffffffu, row, src);
                const uint32_t m = warp_matches(sigs + srow * kSlots, q + (size_t)qi * kSlots, lane);
                if (lane == 0) {
                    const uint32_t pos = atomicAdd(&pcount[qi], 1u);   // beyond cap the entry is dropped: any subset of real rows bounds the k-th best
                    if (pos < cap) pcand[(size_t)qi * cap + pos] = ((uint64_t)(128u - m) << 40) | srow;
                }
            }
        }
    }
}

// (pthr, pkid) = k-th best of a query's probe list, if it found k rows.  The main scan must admit those k rows themselves when
// it reaches them, so the bound it starts from is "at or before (pthr, pkid)" = strictly before (pthr, pkid + 1).
__global__ void adopt_probe_bound_kernel(uint32_t nq, const uint32_t *__restrict__ pthr, const uint64_t *__restrict__ pkid, uint32_t *thr, uint64_t *kth_id) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const uint32_t t = pthr[i];
    uint64_t id = pkid[i];
    if (t == 128 && id == UINT64_MAX) return;   // fewer than k neighbours in the probe: nothing learned
    if (id != UINT64_MAX) id += 1;
    if (t < thr[i] || (t == thr[i] && id < kth_id[i])) { thr[i] = t; kth_id[i] = id; }
}

// exact-selection key for flagged queries: true slot matches of one row, one thread per row
struct JaccardKey {
    static constexpr int kKeyBits = 8;
    static constexpr uint32_t kInvalidKey = 0xFFFFFFFFu;
    __device__ static uint32_t report(uint32_t key, uint32_t flip) { return flip ? flip - key : key; }
    const uint64_t *sigs; const uint64_t *q;
    const uint64_t *qsig;
    __device__ void load_query(uint32_t qi) { qsig = q + (size_t)qi * kSlots; }
    __device__ uint32_t key(uint64_t r) const {
        const uint64_t *rp = sigs + r * kSlots;
        uint32_t m = 0;
        for (int i = 0; i < kSlots; ++i) m += (rp[i] == __ldg(qsig + i));
        return 128u - m;
    }
};

}  // namespace

int jaccard_on_append(ucfp_lane *ctx, ucfp_corpus *c, uint64_t first_row, uint64_t n) {
    if (n == 0) return UCFP_OK;
    uint64_t items = n * kSketchWords;
    sketch_build_kernel<<<(unsigned)((items + 255) / 256), 256, 0, ctx->stream>>>(
        static_cast<const uint64_t *>(c->rows), reinterpret_cast<uint32_t *>(c->mh_sketch),
        reinterpret_cast<uint32_t *>(c->mh_sketch) + sketch_plane_words(c->capacity), first_row, n);
    count_launch(ctx);
    return check_launch("sketch_build");
}

constexpr uint64_t kJaccardExchangeRows[kBoundExchanges] = {1ULL << 12, 1ULL << 14, 1ULL << 16, 1ULL << 19};   // rows seen per rank before exchange e

constexpr size_t kJaccardScanSmemMax = (size_t)kMaxQueriesPerPass * (2 * kSketchWords + 1 + 2) * 4 + 8;

int jaccard_device_init(ucfp_ctx *ctx) {
    UCFP_CUDA_TRY(cudaFuncSetAttribute(compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8192));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(jaccard_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJaccardScanSmemMax));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(jaccard_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kJaccardScanSmemMax + kMaxQueriesPerPass * 4 + 16)));
    int occ = 0;
    UCFP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, jaccard_scan_kernel<false>, kScanThreads, kJaccardScanSmemMax));
    ctx->jac_scan_occ = occ < 1 ? 1 : occ;
    UCFP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, exact_select_kernel<JaccardKey>, 256, 0));
    ctx->jac_exact_occ = occ < 1 ? 1 : (occ > 4 ? 4 : occ);
    return UCFP_OK;
}

int jaccard_scan(ucfp_lane *ctx, ucfp_corpus *c, const uint64_t *q_dev, size_t nq, size_t k, uint64_t *ids_out_dev, uint32_t *m_out_dev) {
    cudaStream_t st = ctx->stream;
    const uint64_t N = c->size;
    UCFP_REQUIRE(k <= 2048, UCFP_E_UNSUPPORTED, "jaccard scan supports k <= 2048 (got %zu)", k);
    UCFP_REQUIRE(N <= kRowMask, UCFP_E_UNSUPPORTED, "corpus too large");
    if (N == 0) {
        size_t tot = nq * k;
        fill_sentinel_u32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ids_out_dev, m_out_dev, tot);
        count_launch(ctx);
        UCFP_TRY(check_launch("fill_sentinel"));
        if (!ctx->xch) return UCFP_OK;   // an empty shard of a group scan still takes part in the bound exchanges below
    }
    uint32_t cap = 4096;
    while (cap < 4 * k) cap <<= 1;
    const uint64_t *sigs = static_cast<const uint64_t *>(c->rows);
    const uint32_t *sketch = reinterpret_cast<const uint32_t *>(c->mh_sketch);
    const uint32_t *sketch1 = sketch + sketch_plane_words(c->capacity);
    const uint64_t *ids = c->id_mode == 1 ? c->ids : nullptr;
    const int occ = ctx->owner->jac_scan_occ;

    for (size_t q0 = 0; q0 < nq; q0 += kMaxQueriesPerPass) {
        const uint32_t nqp = (uint32_t)((nq - q0 < kMaxQueriesPerPass) ? nq - q0 : kMaxQueriesPerPass);
        UCFP_TRY(ctx->qstate.reserve((size_t)nqp * (2 * kSketchWords * 4 + 4 + 8) + 64));
        UCFP_TRY(ctx->cand.reserve(sizeof(uint64_t) * (size_t)cap * nqp));
        UCFP_TRY(ctx->cand_count.reserve(sizeof(uint32_t) * nqp));
        UCFP_TRY(ctx->flags.reserve(sizeof(uint32_t) * (2 * nqp + 1)));
        uint64_t *kth = ctx->qstate.as<uint64_t>();
        uint32_t *qsk = reinterpret_cast<uint32_t *>(kth + nqp);
        uint32_t *thr = qsk + 2 * (size_t)nqp * kSketchWords;   // two query sketch planes precede the bounds
        uint64_t *cand = ctx->cand.as<uint64_t>();
        uint32_t *count = ctx->cand_count.as<uint32_t>();
        uint32_t *flags = ctx->flags.as<uint32_t>();
        const uint64_t *qp = q_dev + q0 * kSlots;
        uint64_t *ids_out = ids_out_dev + q0 * k;
        uint32_t *m_out = m_out_dev + q0 * k;
        SelectState sel{cand, count, thr, 1, kth, flags, cap, flags + nqp, ctx->stats.as<unsigned long long>() + 1, 0, /*small_keys=*/1};

        jaccard_init_kernel<<<(nqp * kSketchWords + 255) / 256, 256, 0, st>>>(qp, nqp, qsk, thr, kth, count, flags);
        if (N == 0) {   // group scan, empty shard: contribute the trivial bound to every exchange
            count_launch(ctx);
            while (ctx->xch->done < kBoundExchanges) UCFP_TRY(exchange_bounds(ctx, sel, nqp));
            ctx->xch->done = 0;
            continue;
        }
        const uint32_t seed = (uint32_t)(N < kSeedRows ? N : kSeedRows);
        jaccard_seed_kernel<<<dim3((seed + 7) / 8, nqp), 256, 0, st>>>(sigs, seed, qp, cand, count, cap);
        count_launch(ctx, 2);
        auto compact = [&](bool final_pass) {
            compact_lists(sel, nqp, (uint32_t)k, ids, c->id_base, final_pass, 128u, ids_out, m_out, st);
            count_launch(ctx, 2);
        };
        compact(seed == N);
        static const bool env_no_probe = getenv("UCFP_JACCARD_NO_PROBE") != nullptr;   // developer switch
        if (!env_no_probe && nqp > 8 && N >= kProbeMinCorpus) {
            const uint64_t pn = (N - seed < kProbeRows) ? N - seed : kProbeRows;
            // probe lists, bounds and flags of their own, in the lane's spill buffer
            const size_t off_cnt = sizeof(uint64_t) * (size_t)cap * nqp, off_kid = off_cnt + sizeof(uint32_t) * 4 * (size_t)nqp;
            UCFP_TRY(ctx->spill.reserve(off_kid + sizeof(uint64_t) * nqp + 64));
            unsigned char *pb = ctx->spill.as<unsigned char>();
            uint64_t *pcand = reinterpret_cast<uint64_t *>(pb);
            uint32_t *pcount = reinterpret_cast<uint32_t *>(pb + off_cnt), *pthr = pcount + nqp, *pflags = pthr + nqp;   // pflags[2 nqp]: overflow, big
            uint64_t *pkid = reinterpret_cast<uint64_t *>(pb + off_kid);
            jaccard_probe_init_kernel<<<(nqp + 255) / 256, 256, 0, st>>>(nqp, pthr, pkid, pcount, pflags);
            const uint64_t ptiles = (pn + kScanThreads - 1) / kScanThreads, pslots = (uint64_t)ctx->sm_count * 4;
            {
                ProfScope ps(ctx, UCFP_PROF_JACCARD_SCAN, 256.0 * (double)pn * nqp);   // a quarter of the slots
                jaccard_probe_kernel<<<(unsigned)(ptiles < pslots ? ptiles : pslots), kScanThreads, (size_t)nqp * kProbeWords * 4, st>>>(
                    sketch, sigs, seed, pn, qp, qsk, nqp, pcand, pcount, cap);
            }
            SelectState probe{pcand, pcount, pthr, 1, pkid, pflags, cap, pflags + nqp, nullptr, 0, /*small_keys=*/1};
            compact_lists(probe, nqp, (uint32_t)k, ids, c->id_base, false, 128u, nullptr, nullptr, st);
            adopt_probe_bound_kernel<<<(nqp + 255) / 256, 256, 0, st>>>(nqp, pthr, pkid, thr, kth);
            count_launch(ctx, 5);
        }
        // Chunks scanned under a loose bound cost ~8x more per row (no early exit, every chance collision is verified), and the
        // bound only tightens between chunks: x4 steps reach a useful bound after ~340 K rows of config 3 instead of ~590 K.
        static const long env_growth = getenv("UCFP_JACCARD_GROWTH") ? atol(getenv("UCFP_JACCARD_GROWTH")) : 0;   // developer knob
        const uint64_t growth = env_growth > 1 ? (uint64_t)env_growth : (nqp <= 8 ? 32 : 4);
        uint64_t pos = seed, chunk = (uint64_t)seed * growth;
        const size_t smem = (size_t)nqp * (2 * kSketchWords + 1 + 2) * 4 + 8;
        while (pos < N) {
            uint64_t n = (N - pos < chunk) ? N - pos : chunk;
            uint64_t ntiles = (n + kScanThreads - 1) / kScanThreads;
            const uint64_t slots_full = (uint64_t)ctx->sm_count * occ;
            uint64_t grid = slots_full, q_groups = 1;
            if (ntiles < slots_full) {   // small chunk: split the queries so that every SM gets work
                q_groups = (slots_full + ntiles - 1) / ntiles;
                const uint64_t max_groups = (nqp + 7) / 8;   // at least 8 queries per group
                if (q_groups > max_groups) q_groups = max_groups;
                grid = ntiles * q_groups;
            }
            {
                ProfScope ps(ctx, UCFP_PROF_JACCARD_SCAN, 1024.0 * (double)n * nqp);
                jaccard_scan_kernel<false><<<(unsigned)grid, kScanThreads, smem, st>>>(sketch, sketch1, sigs, ids, c->id_base, pos, n, qp, qsk, nqp,
                                                                                      (uint32_t)q_groups, sel);
            }
            count_launch(ctx);
            pos += n;
            compact(pos == N);
            // group scan: the loose-bound phase is where a small shard loses its time (no early exit, every chance collision is
            // verified); the other shards' bounds end it after a few thousand rows per rank instead of a few hundred thousand
            if (ctx->xch && pos < N)
                while (ctx->xch->done < kBoundExchanges && pos >= kJaccardExchangeRows[ctx->xch->done]) UCFP_TRY(exchange_bounds(ctx, sel, nqp));
            chunk = chunk * growth < kMaxChunkRows ? chunk * growth : kMaxChunkRows;
        }
        if (ctx->xch) {
            while (ctx->xch->done < kBoundExchanges) UCFP_TRY(exchange_bounds(ctx, sel, nqp));
            ctx->xch->done = 0;
        }
        UCFP_TRY(check_launch("jaccard scan"));
        // Queries whose candidate list overflowed (device-side decisions, no host sync): up to kRescanRounds passes over the sketches
        // for the flagged queries alone under the bound their truncated lists produced; the launches return at once when nothing
        // is flagged.  What is still flagged afterwards goes to the exact multi-pass selection.
        UCFP_TRY(stats_add_flags(ctx, flags, nqp));
        for (int round = 0; round < rescan_rounds(); ++round) {
            jaccard_scan_kernel<true><<<(unsigned)(ctx->sm_count * occ), kScanThreads, smem + (size_t)nqp * 4 + 16, st>>>(
                sketch, sketch1, sigs, ids, c->id_base, 0, N, qp, qsk, nqp, 1u, sel);
            compact_rescanned(sel, nqp, (uint32_t)k, ids, c->id_base, 128u, ids_out, m_out, st);
            count_launch(ctx, 2);
        }
        UCFP_TRY(exact_select_fallback(ctx, c, ctx->owner->jac_exact_occ, JaccardKey{sigs, qp, nullptr}, flags, nqp, (uint32_t)k, 128u, ids_out, m_out));
    }
    return UCFP_OK;
}

}  // namespace ucfp
