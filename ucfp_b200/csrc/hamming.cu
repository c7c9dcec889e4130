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
//   Batches of >= kMmaMinQueries queries run the large chunks on the int8 tensor pipe instead
//   (hamming_mma_scan_kernel below): the POPC pipe (16 lanes/clk/SM) caps the loop above at ~14
//   pairs/clk/SM, the tensor-core form filters ~96 pairs/clk/SM; same admission rule, same lists.
// A list that overflows its capacity (adversarial duplicates with descending ids) is
// flagged; flagged queries are recomputed by the exact multi-pass selection in
// topk_select.cuh (histogram of distances + radix select on ids), so the result is
// exact for every input.
#include <cooperative_groups.h>
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace ucfp {

namespace {

#include "topk_select.cuh"
#include "sm100_ptx.cuh"

constexpr int kScanThreads = 256;
constexpr int kCodesPerThread = 8;                 // 4 x LDG.128 in flight per thread
constexpr int kTileCodes = kScanThreads * kCodesPerThread;
constexpr uint32_t kSeedRows = 1024;               // a multiple of 512: chunk starts stay aligned to the tensor scan's stage images; 2k for large k
constexpr uint32_t kMaxQueriesPerPass = 1024;      // POPC scan: 16 B/query of shared memory; tensor scan: 8 resident 128-query tiles
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
                cand_append(cand, count, cap, q, ((uint64_t)d[c] << 40) | r);
            }
        }
    }
}

// Scans rows [row0, row0 + nrows) of the corpus against the nq staged queries.  kRescan: against the FLAGGED queries only (a
// re-scan round after a candidate list overflowed, topk_select.cuh); with no flag set every CTA returns at once.
template <bool kRescan>
__global__ void __launch_bounds__(kScanThreads, 4)
hamming_scan_kernel(const uint64_t *__restrict__ codes, const uint64_t *__restrict__ ids, uint64_t id_base,
                    uint64_t row0, uint64_t nrows, const QSlot *__restrict__ slots, const uint64_t *__restrict__ kth_id,
                    uint32_t nq, uint32_t q_groups, uint64_t *cand, uint32_t *count, uint32_t cap, const uint32_t *__restrict__ flags) {
    extern __shared__ uint4 sq[];  // query slots of this CTA's group
    // Small chunks (the first few of a batch, where the admission bounds are still loose and the cold path runs
    // often) have fewer tiles than the GPU has CTA slots: the queries are then split into q_groups groups and the
    // grid is tiles x groups, one (tile, group) pair per CTA.  Large chunks use q_groups == 1 and a persistent grid.
    const uint32_t grp = q_groups > 1 ? blockIdx.x % q_groups : 0;
    const uint32_t q_lo = (uint32_t)((uint64_t)nq * grp / q_groups), q_hi = (uint32_t)((uint64_t)nq * (grp + 1) / q_groups);
    uint32_t nq_mine = q_hi - q_lo;
    if constexpr (kRescan) {   // slot.pad carries the query's index
        __shared__ uint32_t s_listed;
        if (threadIdx.x < 32) {
            const uint32_t listed = list_flagged_queries(flags, nq, [&](uint32_t pos, uint32_t q) {
                const QSlot s = slots[q];
                sq[pos] = make_uint4(s.lo, s.hi, s.thr, q);
            });
            if (threadIdx.x == 0) s_listed = listed;
        }
        __syncthreads();
        nq_mine = s_listed;
        if (nq_mine == 0) return;
    } else {
        for (uint32_t i = threadIdx.x; i < nq_mine; i += kScanThreads) sq[i] = reinterpret_cast<const uint4 *>(slots)[q_lo + i];
        __syncthreads();
    }

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
                                  make_uint4(lo[4], hi[4], lo[5], hi[5]), make_uint4(lo[6], hi[6], lo[7], hi[7]), s, kRescan ? s.w : q_lo + q,
                                  tile_row, row_end, ids, id_base, kth_id, cand, count, cap);
        }
    }
}

// ---- batched scan on the int8 tensor pipe ---------------------------------------------------------
// A Hamming distance matrix is a binary GEMM.  With bits mapped to +-1 (set -> +1, clear -> -1) the dot product of a
// query and a code is  x = 64 - 2 * dist.  tcgen05.mma kind::i8 (s8 x s8 -> s32 in TMEM) computes 128 queries x 256
// operand rows per instruction pair (K = 64 = 2 x 32).  Each operand row carries TWO codes a, b as  -a_k + 64 * b_k
// (values +-63, +-65 fit s8), so one accumulator is  D = -x_a + 64 * x_b,  |D| <= 4160, and holds both distances:
//     x_a >= tau   <=>   the low 7 bits of D, read as a signed number, are <= -tau     (x_a = 64 and x_a = -64 alias;
//                                                                                       the latter is dist 64: a harmless false positive)
//     x_b >= tau   <=>   D >= 64 * tau - 64                                             (x_a = -64 again the only false positive)
// with tau = 64 - 2 * thr.  The epilogue never decodes distances: per thread (= one query) it reads 64 accumulators
// as 32 registers of s16 pairs (tcgen05.ld ... pack::16b) and keeps a per-halfword signed max of D (the x_b test) and
// min of D * 512 (x_a's 7 bits at the top of each halfword) with VIMNMX3.S16x2: 1 multiply + 1 min/max lane-op per
// register = per four (query, code) pairs, and the per-query bounds come ready-made from shared memory.
// Only when a bound is crossed does that lane take the cold path: it parks its 128-code strip for recheck_parked_kernel (or, when the
// CTA's queue is full, decodes the distances from its registers) under the same admission rule as hamming_survivors.  Queries stay
// resident in shared memory as eight 128-row A tiles.  Template parameters:
//   kPreExpanded  operand rows copied ready-made from corpus->ham_ops by TMA bulk copies (64-640 queries) / expanded from the codes
//                 by producer warps in the kernel (641-1024 queries, and any corpus without stage images);
//   kStagger      epilogue as two groups of eight warps, group g serving accumulator stage g (every second item), 128 columns per
//                 warp as two 64-column strips in two register images -- the product path of both forms (20 warps at most, so that
//                 96 registers are available) / all sixteen warps in lock-step on the same item, 64 columns each: round 1's schedule,
//                 kept for batches of 64-640 queries on corpora without stage images (three producer warps do not keep up there) and
//                 behind UCFP_HAMMING_STAGGER=0 / UCFP_HAMMING_STAGGER_EXP=0.
constexpr int kMmaQTile = 128;                       // UMMA M: queries per accumulator tile (TMEM lanes)
constexpr int kMmaRows = 256;                        // UMMA N: operand rows per stage = 512 codes (TMEM columns)
constexpr int kMmaTileCodes = 2 * kMmaRows;
constexpr int kMmaRowBytes = 128;                    // one 128B-swizzle row; bytes 0..63 hold the K = 64 elements
constexpr int kMmaQBytes = kMmaQTile * kMmaRowBytes; // 16 KiB per query tile
constexpr int kMmaCBytes = kMmaRows * kMmaRowBytes;  // 32 KiB per code stage
constexpr int kMmaStages = 2;                        // in-kernel expansion: 2 x 32 KiB (128-byte rows, SWIZZLE_128B)
constexpr int kMmaImgBytes = kMmaRows * 64;          // pre-built stage image: 16 KiB (64-byte rows, SWIZZLE_64B)
constexpr int kMmaMaxImgStages = 8;
constexpr int kMmaExpWarps = 4, kMmaEpiWarps = 16;
constexpr int kMmaThreads = 32 * (1 + kMmaExpWarps + kMmaEpiWarps);
// two-group epilogue (two 32-register accumulator images per epilogue thread): at most 20 warps, so that ptxas may use 96 registers.
// Stage-image form: MMA issuer + one TMA thread + 16 epilogue warps; expansion form: MMA issuer + THREE producer warps + 16.
constexpr int kMmaStaggerExpWarps = 3;
constexpr int kMmaStaggerThreads = 32 * (2 + kMmaEpiWarps), kMmaStaggerExpThreads = 32 * (1 + kMmaStaggerExpWarps + kMmaEpiWarps);
constexpr int kMmaColsPerWarp = kMmaRows / (kMmaEpiWarps / 4);
constexpr uint32_t kMmaMaxQueries = 1024;
constexpr uint32_t kMmaMinQueries = 64;              // measured crossover: the POPC scan costs 0.24 ms per query and 1 B rows, the tensor scan >= 15 ms per batch
constexpr uint64_t kMmaMinChunkRows = 1ULL << 16;    // smaller chunks (fewer tiles than SMs, very loose bounds) stay on the POPC scan
constexpr size_t kMmaSmem = (size_t)(kMmaMaxQueries / kMmaQTile) * kMmaQBytes + kMmaStages * kMmaCBytes + kMmaMaxQueries * (16 + 8 + 4 + 4) + 256 + 1024;
static_assert(kMmaMaxQueries <= kMaxQueriesPerPass || kMaxQueriesPerPass <= kMmaMaxQueries, "");
static_assert(kMmaColsPerWarp == 64, "one packed tcgen05.ld per warp and accumulator tile");

// 4 bits -> 4 bytes 0/1
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }
// query row: 64 elements +-1 in the first four 16-byte chunks of row r of a 128B-swizzled K-major tile
__device__ __forceinline__ void mma_store_query_row(unsigned char *tile, uint32_t r, uint32_t lo, uint32_t hi, bool valid) {
    unsigned char *row = tile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t h = (c < 2 ? lo : hi) >> ((c & 1) * 16);
        uint4 w;   // bit set -> 0x01, clear -> 0xFF
        w.x = ~(spread4(h & 15) * 0xFEu); w.y = ~(spread4((h >> 4) & 15) * 0xFEu);
        w.z = ~(spread4((h >> 8) & 15) * 0xFEu); w.w = ~(spread4((h >> 12) & 15) * 0xFEu);
        if (!valid) w = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(row + ((c ^ (r & 7)) << 4)) = w;
    }
}
// operand row of two codes: element k = -a_k + 64 b_k = 0xC1 ^ (abit * 0x7E) ^ (bbit * 0x80)
__device__ __forceinline__ uint32_t mma_pack4(uint32_t na, uint32_t nb) { return 0xC1C1C1C1u ^ (spread4(na) * 0x7Eu) ^ (spread4(nb) << 7); }
__device__ __forceinline__ void mma_store_code_row(unsigned char *tile, uint32_t r, uint64_t a, uint64_t b) {
    unsigned char *row = tile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t ha = (uint32_t)(a >> (16 * c)) & 0xFFFFu, hb = (uint32_t)(b >> (16 * c)) & 0xFFFFu;
        uint4 w;
        w.x = mma_pack4(ha & 15, hb & 15); w.y = mma_pack4((ha >> 4) & 15, (hb >> 4) & 15);
        w.z = mma_pack4((ha >> 8) & 15, (hb >> 8) & 15); w.w = mma_pack4(ha >> 12, hb >> 12);
        *reinterpret_cast<uint4 *>(row + ((c ^ (r & 7)) << 4)) = w;
    }
}

// Operand rows built once at append time (corpus->ham_ops): row g = codes 2g and 2g + 1, 64 bytes.  They are stored as the
// shared-memory IMAGE of the scan's stage tiles -- 256 rows = 16 KiB per tile, 64B-swizzled (chunk c of row r at chunk
// c ^ ((r >> 1) & 3)) -- so that the scan's producer is one TMA bulk copy per stage.
__global__ void hamming_ops_kernel(const uint64_t *__restrict__ codes, uint64_t row_lo, uint64_t row_hi, uint64_t size, uint4 *__restrict__ ops) {
    const uint64_t g = row_lo / 2 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * g >= row_hi) return;
    const uint64_t a = codes[2 * g], b = 2 * g + 1 < size ? codes[2 * g + 1] : 0;
    const uint32_t r = (uint32_t)(g % kMmaRows);
    uint4 *row = ops + 4 * g;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t ha = (uint32_t)(a >> (16 * c)) & 0xFFFFu, hb = (uint32_t)(b >> (16 * c)) & 0xFFFFu;
        row[c ^ ((r >> 1) & 3)] = make_uint4(mma_pack4(ha & 15, hb & 15), mma_pack4((ha >> 4) & 15, (hb >> 4) & 15),
                                             mma_pack4((ha >> 8) & 15, (hb >> 8) & 15), mma_pack4(ha >> 12, hb >> 12));
    }
}

struct MmaScanArgs {
    const uint64_t *codes; const uint4 *ops; const uint64_t *ids; uint64_t id_base;   // ops: pre-built stage images or null
    uint64_t row0, row_end;                  // rows [row0, row_end), row0 even
    const QSlot *slots; const uint64_t *kth_id; uint32_t nq;
    uint64_t *cand; uint32_t *count; uint32_t cap;
    uint32_t wait_flags;                     // experiments only (hamming_experiments.cuh)
    // A lane whose hot test fired does not settle its 128 rows inside the scan: it parks {query, first row} in a per-CTA queue in
    // global memory (slot taken with a SHARED-memory atomic, store fire-and-forget) and recheck_parked_kernel settles the
    // parked strips after the launch with ordinary popcounts.  Settling in place costs a lone lane ~500 instructions while the
    // other fifteen epilogue warps and the MMA issuer wait two accumulator stages later: in a 125M-row shard that was 0.74 ms
    // of 5.4 ms (ncu launch list, profiles/r02_launches_125m.md).
    uint4 *spill; uint32_t *spill_count; uint32_t spill_cap;
};

constexpr uint32_t kRecheckSlices = 8;       // CTAs of recheck_parked_kernel per queue
constexpr uint32_t kSpillCap = 4096;         // strips per CTA and launch (64 KiB); beyond it a lane settles in place as before

// Cold path of one lane (= one query) whose bounds were crossed somewhere in its 64 accumulators.  It runs AFTER the warp
// has handed the TMEM stage back, from the packed register image alone (|D| <= 4160 fits the 16 bits that were loaded),
// so a fire delays one warp, not the CTA's MMA pipeline.  Nothing waits on global memory except the list append: query
// slots and k-th ids sit in shared memory and both distances are decoded from D = -x_a + 64 x_b  (u = 64 - x_a is D's
// low 7 bits ^ 64; x_b = (D + x_a) / 64).  Only u == 0, where x_a = 64 and -64 alias, reads the two codes.
template <int NREG>   // NREG packed registers = 2 NREG accumulator columns = 4 NREG codes starting at first_row
__device__ __forceinline__ void hamming_mma_settle(const uint32_t (&p)[NREG], uint64_t first_row, uint32_t thr_hot, uint32_t q,
                                                   const MmaScanArgs &A, const uint4 *s_q, const uint64_t *s_kid) {
    const int32_t hi_bound = 64 * (63 - 2 * (int32_t)thr_hot);
    uint32_t m_even = 0, m_odd = 0;   // bit c: column 2c / 2c + 1 can hold an admissible pair
#pragma unroll
    for (int c = 0; c < NREG; ++c) {
        const int32_t de = (int32_t)(int16_t)(p[c] & 0xFFFFu), dq = (int32_t)p[c] >> 16;
        if (((((uint32_t)de ^ 64u) & 127u) <= 2 * thr_hot) | (de >= hi_bound)) m_even |= 1u << c;
        if (((((uint32_t)dq ^ 64u) & 127u) <= 2 * thr_hot) | (dq >= hi_bound)) m_odd |= 1u << c;
    }
    const uint4 s = s_q[q];          // {lo, hi, thr, -}
    const uint32_t thr = s.z;
    const uint64_t kid = s_kid[q];
    while (m_even | m_odd) {
        const bool odd = m_even == 0;
        uint32_t &m = odd ? m_odd : m_even;
        const int c = __ffs(m) - 1;
        m &= m - 1;
        uint32_t reg = 0;
#pragma unroll
        for (int j = 0; j < NREG; ++j) reg = (c == j) ? p[j] : reg;   // register select: no local-memory copy of p[]
        const int32_t D = odd ? (int32_t)reg >> 16 : (int32_t)(int16_t)(reg & 0xFFFFu);
        const uint64_t r0 = first_row + 2 * (2 * c + (odd ? 1 : 0));
        const uint32_t u = ((uint32_t)D ^ 64u) & 127u;
        uint32_t d[2];
        if (u != 0) {
            const int32_t xa = 64 - (int32_t)u, xb = (D + xa) >> 6;
            d[0] = u >> 1; d[1] = (uint32_t)(64 - xb) >> 1;
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t code = r0 + h < A.row_end ? A.codes[r0 + h] : 0;
                d[h] = __popc((uint32_t)code ^ s.x) + __popc((uint32_t)(code >> 32) ^ s.y);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint64_t r = r0 + h;
            if (r >= A.row_end || d[h] > thr) continue;
            const uint64_t id = A.ids ? (d[h] == thr ? A.ids[r] : 0) : A.id_base + r;
            if (d[h] < thr || id < kid) {
                cand_append(A.cand, A.count, A.cap, q, ((uint64_t)d[h] << 40) | r);
            }
        }
    }
}

// After a tensor-scan launch: one warp per parked strip {query, first row} settles its 128 rows (lane l: rows 4l .. 4l + 3) under
// the same admission rule as hamming_mma_settle -- d < thr, or d == thr and id below the k-th result's -- and appends to the
// query's candidate list (same counters, same overflow rule).  The bounds are the ones the scan launch itself used: only the
// compaction that follows changes them.
__global__ void __launch_bounds__(256) recheck_parked_kernel(const uint4 *__restrict__ spill, const uint32_t *__restrict__ spill_count, uint32_t spill_cap,
                                                              const uint64_t *__restrict__ codes, const uint64_t *__restrict__ ids, uint64_t id_base, uint64_t row_end,
                                                              const QSlot *__restrict__ slots, const uint64_t *__restrict__ kth_id,
                                                              uint64_t *cand, uint32_t *count, uint32_t cap) {
    const uint32_t n = min(spill_count[blockIdx.x], spill_cap);
    const uint4 *mine = spill + (size_t)blockIdx.x * spill_cap;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // gridDim.y CTAs share a queue: the chain entry -> query slot -> codes -> append is all latency, so it wants many warps
    for (uint32_t i = blockIdx.y * (blockDim.x >> 5) + warp; i < n; i += gridDim.y * (blockDim.x >> 5)) {
        const uint4 e = mine[i];
        const uint32_t q = e.x;
        const uint64_t r0 = ((uint64_t)e.w << 32 | e.z) + 4 * lane;   // even first row: 16-byte aligned pairs
        const QSlot s = slots[q];
        const uint64_t kid = kth_id[q];
        uint64_t code[4] = {0, 0, 0, 0};
        if (r0 + 3 < row_end) {
            const uint4 a = *reinterpret_cast<const uint4 *>(codes + r0), b = *reinterpret_cast<const uint4 *>(codes + r0 + 2);
            code[0] = (uint64_t)a.y << 32 | a.x; code[1] = (uint64_t)a.w << 32 | a.z;
            code[2] = (uint64_t)b.y << 32 | b.x; code[3] = (uint64_t)b.w << 32 | b.z;
        } else {
#pragma unroll
            for (int h = 0; h < 4; ++h) if (r0 + h < row_end) code[h] = codes[r0 + h];
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint64_t r = r0 + h;
            const uint32_t d = __popc((uint32_t)code[h] ^ s.lo) + __popc((uint32_t)(code[h] >> 32) ^ s.hi);
            if (r >= row_end || d > s.thr) continue;
            const uint64_t id = ids ? (d == s.thr ? ids[r] : 0) : id_base + r;
            if (d < s.thr || id < kid) {
                cand_append(cand, count, cap, q, ((uint64_t)d << 40) | r);
            }
        }
    }
}

#ifdef UCFP_HAMMING_DIAG   // developer build: A.wait_flags bit 0 skips the min/max test, bit 1 also the TMEM loads (what is left is the MMA pipeline and its hand-overs)
#define UCFP_DIAG_NO_TEST(A) ((A).wait_flags & 3u)
#define UCFP_DIAG_NO_LD(A) ((A).wait_flags & 2u)
#else
#define UCFP_DIAG_NO_TEST(A) false
#define UCFP_DIAG_NO_LD(A) false
#endif
[[maybe_unused]] constexpr uint32_t kMmaNeverHiPk = 0x7FFF7FFFu;   // hi16 - 1 (both halfwords) of a query that can never fire: no accumulator crosses it
// The hot test of one 64-column strip: per-halfword signed max of D and min of D << 9 (VIMNMX3.S16x2: two columns per lane-op)
// against the query's two bounds; true when some halfword exceeded hi16 - 1 or fell below lo16 + 1.
// The low field's view of a packed register: the 7 low bits of each halfword moved to its top.  p * 512 compiles to IMAD.SHL (FMA
// pipe).  A rotation by 9 (funnel shift SHF.L.W, ALU pipe; the wrapped-in bits land in the 9 junk bits the bounds already mask) is
// equivalent and was measured because a mixed microbenchmark loop pairs VIMNMX3.S16x2 with ALU-pipe logic at one instruction per clock
// and with IMAD.SHL at 1.52 (profiles/r02_pipe_rates.txt) -- in the kernel it is slower: 1 024 queries x 1 B codes 35.56 ms with the
// multiply, 43.40 with the rotation, 37.10 alternating the two (scripts/r2/gpu_step38.sh).  -DUCFP_HAMMING_SHIFT=1|2 rebuilds them.
#ifndef UCFP_HAMMING_SHIFT
#define UCFP_HAMMING_SHIFT 0
#endif
template <int kWhich>
__device__ __forceinline__ uint32_t hamming_mma_low_view(uint32_t p) {
    if (UCFP_HAMMING_SHIFT == 0 || (UCFP_HAMMING_SHIFT == 2 && (kWhich & 1))) return p * 512u;
    return __funnelshift_l(p, p, 9);
}
template <bool kMinimalOps>
__device__ __forceinline__ bool hamming_mma_strip_test(const uint32_t (&p)[32], uint32_t hi_pk, uint32_t lo_pk) {
    uint32_t mx[4], mn[4];   // four independent chains per test (latency)
#pragma unroll
    for (int j = 0; j < 4; ++j) { mx[j] = p[j]; mn[j] = hamming_mma_low_view<0>(p[j]); }
    if constexpr (kMinimalOps) {
        // 32 registers + the bound = 33 inputs per test: sixteen 3-input operations is the minimum.  Two chains of nine and two of seven
        // registers, so that no chain ends in a 2-input operation: 14 + 2 VIMNMX3 per test instead of 16 + 2.  Measured on one box
        // (scripts/r2/gpu_step35.sh): expansion form, 1 024 queries x 1 B codes 35.23 -> 34.36 ms; the image form at 512 queries is 2.5 %
        // SLOWER with it (5.17 -> 5.30 ms per 250 M codes) and keeps the even chains below.
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // chains 0 and 1: registers 4..11 and 12..19
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = 4 + 8 * j + 2 * i;
                mx[j] = __vimax3_s16x2(mx[j], p[c], p[c + 1]);
                mn[j] = __vimin3_s16x2(mn[j], hamming_mma_low_view<0>(p[c]), hamming_mma_low_view<1>(p[c + 1]));
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {   // chains 2 and 3: registers 20..25 and 26..31
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = 20 + 6 * j + 2 * i;
                mx[2 + j] = __vimax3_s16x2(mx[2 + j], p[c], p[c + 1]);
                mn[2 + j] = __vimin3_s16x2(mn[2 + j], hamming_mma_low_view<0>(p[c]), hamming_mma_low_view<1>(p[c + 1]));
            }
        }
    } else {
#pragma unroll
        for (int c = 4; c < 28; c += 8)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                mx[j] = __vimax3_s16x2(mx[j], p[c + j], p[c + 4 + j]);
                mn[j] = __vimin3_s16x2(mn[j], hamming_mma_low_view<0>(p[c + j]), hamming_mma_low_view<1>(p[c + 4 + j]));
            }
#pragma unroll
        for (int j = 0; j < 4; ++j) { mx[j] = __vmaxs2(mx[j], p[28 + j]); mn[j] = __vmins2(mn[j], hamming_mma_low_view<0>(p[28 + j])); }
    }
    const uint32_t m2 = __vimax3_s16x2(__vimax3_s16x2(mx[0], mx[1], mx[2]), mx[3], hi_pk);
    const uint32_t n2 = __vimin3_s16x2(__vimin3_s16x2(mn[0], mn[1], mn[2]), mn[3], lo_pk);
    return (m2 != hi_pk) | (n2 != lo_pk);
}

// A fired strip is parked for recheck_parked_kernel, or settled in place from the register image when the CTA's queue is full.
__device__ __forceinline__ void hamming_mma_fire(const uint32_t (&p)[32], uint64_t first_row, uint32_t q, const MmaScanArgs &A,
                                                 uint32_t *s_spill_n, const uint4 *s_q, const uint64_t *s_kid) {
    const uint32_t slot = A.spill ? atomicAdd(s_spill_n, 1u) : 0xFFFFFFFFu;
    if (slot < A.spill_cap) {   // park the strip: nothing waits for this store
        A.spill[(size_t)blockIdx.x * A.spill_cap + slot] = make_uint4(q, 0u, (uint32_t)first_row, (uint32_t)(first_row >> 32));
    } else {   // queue full (or switched off): settle in place; the hot test's bound is thr - 1 under implicit ids
        const uint32_t thr = s_q[q].z;
        hamming_mma_settle<32>(p, first_row, (A.ids == nullptr && s_kid[q] < A.id_base + A.row0) ? (thr == 0 ? 0u : thr - 1) : thr, q, A, s_q, s_kid);
    }
}


template <bool kPreExpanded, bool kStagger>
__global__ void __launch_bounds__(kStagger ? (kPreExpanded ? kMmaStaggerThreads : kMmaStaggerExpThreads) : kMmaThreads, 1)
hamming_mma_scan_kernel(const __grid_constant__ MmaScanArgs A) {
    constexpr int kProdWarps = kStagger ? (kPreExpanded ? 1 : kMmaStaggerExpWarps) : kMmaExpWarps;   // warps between the MMA issuer and the epilogue warps
    constexpr int kProdThreads = 32 * kProdWarps, kProdRows = (kMmaRows + kProdThreads - 1) / kProdThreads;   // expansion: operand rows per producer thread and stage
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q_tiles = (A.nq + kMmaQTile - 1) / kMmaQTile;
    unsigned char *sQ = smem;                                                          // [q_tiles][16 KiB] query tiles
    // operand-row stages: expansion path 2 x 32 KiB after all eight query-tile slots; image path 16 KiB each, starting right
    // after the query tiles in use (4 stages at 1024 queries, 8 at <= 512)
    unsigned char *sC = kPreExpanded ? smem + (size_t)q_tiles * kMmaQBytes : smem + (size_t)(kMmaMaxQueries / kMmaQTile) * kMmaQBytes;
    const uint32_t n_stages = kPreExpanded ? min((uint32_t)kMmaMaxImgStages, 12u - q_tiles) : (uint32_t)kMmaStages;
    const uint32_t stage_bytes = kPreExpanded ? kMmaImgBytes : kMmaCBytes;
    uint4 *s_q = reinterpret_cast<uint4 *>(smem + (size_t)(kMmaMaxQueries / kMmaQTile) * kMmaQBytes + kMmaStages * kMmaCBytes);   // [1024] query slots {lo, hi, thr, -}
    uint64_t *s_kid = reinterpret_cast<uint64_t *>(s_q + kMmaMaxQueries);              // [1024] id of the current k-th result
    uint2 *s_bnd = reinterpret_cast<uint2 *>(s_kid + kMmaMaxQueries);                  // [1024] the hot test's two s16 bounds, see below
    uint64_t *cfull = reinterpret_cast<uint64_t *>(s_bnd + kMmaMaxQueries);
    uint64_t *cempty = cfull + kMmaMaxImgStages, *tfull = cempty + kMmaMaxImgStages, *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    uint32_t *s_spill_n = tmem_slot + 1;   // entries this CTA has parked in its global queue
    const uint32_t n_tiles = (uint32_t)((A.row_end - A.row0 + kMmaTileCodes - 1) / kMmaTileCodes);

    for (uint32_t q = threadIdx.x; q < q_tiles * kMmaQTile; q += blockDim.x) {
        const QSlot s = q < A.nq ? A.slots[q] : QSlot{0, 0, 0, 0};
        mma_store_query_row(sQ + (q / kMmaQTile) * kMmaQBytes, q % kMmaQTile, s.lo, s.hi, q < A.nq);
        const uint64_t kid = q < A.nq ? A.kth_id[q] : 0;
        s_q[q] = make_uint4(s.lo, s.hi, s.thr, 0);
        s_kid[q] = kid;
        // Implicit ids grow with the row index; when the k-th result's id precedes every row of this launch (always, in a
        // single-corpus scan: it is an earlier row; in a group scan the bound may come from a shard with larger ids) a tie
        // at distance thr can never be admitted: the hot test may use thr - 1 (about 4x fewer trips through the cold path).
        uint32_t hot = s.thr;
        if (A.ids == nullptr && kid < A.id_base + A.row0) hot = s.thr == 0 ? 0xFFFFFFFFu : s.thr - 1;
        if (q >= A.nq) hot = 0xFFFFFFFFu;
        // The hot test works on the 16-bit image of an accumulator: field y = the value itself, field x = its low 7 bits moved to
        // the top of the halfword by * 512 (the upper halfword of a register then carries < 512 of junk from the lower one, hence
        // the | 0x1FF).  It fires when  max(y) >= hi16  or  min(x) <= lo16;  stored as hi16 - 1 and lo16 + 1 so that both become
        // "max/min against the bound changes the bound".  never: neither bound can be crossed by |D| <= 4160; thr >= 64: always.
        const int32_t tau = 64 - 2 * (int32_t)hot;
        int32_t hi16 = 64 * (tau - 1), lo16 = (-tau * 512) | 0x1FF;
        if (hot == 0xFFFFFFFFu) { hi16 = 0x8000; lo16 = -0x8001; }   // bounds 0x7FFF / -0x8000: no halfword exceeds the one or falls below the other
                                                                      // (-0x7FFF + 1 left a hole: a low field of -64, i.e. an exact duplicate in an even row, still fired)
        else if (hot >= 64) { hi16 = -0x7FFF; lo16 = 0; }
        const uint32_t hi1 = (uint16_t)(hi16 - 1), lo1 = (uint16_t)(lo16 + 1);
        s_bnd[q] = make_uint2(hi1 << 16 | hi1, lo1 << 16 | lo1);   // both halfwords
    }
    if (threadIdx.x == 0) {
        *s_spill_n = 0;
        for (uint32_t s = 0; s < n_stages; ++s) { mbar_init(&cfull[s], kPreExpanded ? 1 : kProdThreads); mbar_init(&cempty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kStagger ? kMmaEpiWarps / 2 : kMmaEpiWarps); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc_512(tmem_slot);   // two accumulator stages of 256 columns
    fence_proxy_async_smem();                   // the query tiles were written through the generic proxy
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== MMA issuer: D[128 queries x 256 rows] = Q_tile (K-major, 64 x s8) * rows^T, two K = 32 steps =====
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kMmaRows >> 3) << 17) | ((uint32_t)(kMmaQTile >> 4) << 24);
        uint32_t it = 0, acc_it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
            mbar_wait_sleep(&cfull[s], ph);
            tcgen05_fence_after();
            const uint64_t bdesc = kPreExpanded ? umma_desc_sw64(smem_u32(sC + s * stage_bytes)) : umma_desc_sw128(smem_u32(sC + s * stage_bytes));
            for (uint32_t mt = 0; mt < q_tiles; ++mt, ++acc_it) {
                const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                mbar_wait_sleep(&tempty[as], aph ^ 1);
                tcgen05_fence_after();
                if (lane == 0) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(sQ + mt * kMmaQBytes));
                    umma_i8(tmem_base + as * kMmaRows, adesc, bdesc, idesc, 0u);
                    umma_i8(tmem_base + as * kMmaRows, adesc + 2, bdesc + 2, idesc, 1u);   // +32 bytes of K
                    umma_commit(&tfull[as]);
                }
                __syncwarp();
            }
            if (lane == 0) umma_commit(&cempty[s]);   // frees the operand stage when its MMAs retire
            __syncwarp();
        }
    } else if (warp <= kProdWarps) {
        // ===== producers: thread t fills operand rows t, t + kProdThreads, ... of a stage =====
        const uint32_t t = threadIdx.x - 32;
        uint32_t it = 0;
        if (kPreExpanded) {
            // stage images come ready-made from HBM (corpus->ham_ops): one elected thread, one 16 KiB TMA bulk copy per stage,
            // as many in flight as there are stages
            if (threadIdx.x == 32) {
                const unsigned char *src = reinterpret_cast<const unsigned char *>(A.ops) + (A.row0 / kMmaTileCodes) * (uint64_t)kMmaImgBytes;
                for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                    const uint32_t s = it % n_stages, ph = (it / n_stages) & 1;
                    mbar_wait_sleep(&cempty[s], ph ^ 1);
                    mbar_expect_tx(&cfull[s], kMmaImgBytes);
                    tma_bulk_g2s(sC + s * kMmaImgBytes, src + (uint64_t)tile * kMmaImgBytes, kMmaImgBytes, &cfull[s]);
                }
            }
        } else {
            // thread t expands codes 2r, 2r + 1 (16-byte loads) for the operand rows r = t, t + kProdThreads, ... of a stage
            auto load_tile = [&](uint32_t tile, uint64_t (&c)[2 * kProdRows]) {
                const uint64_t base = A.row0 + (uint64_t)tile * kMmaTileCodes;
#pragma unroll
                for (int j = 0; j < kProdRows; ++j) {   // rows past the end become code 0 and are rejected by the cold path's range check
                    const uint64_t r0 = base + 2 * (t + kProdThreads * j);
                    if (kMmaRows % kProdThreads != 0 && t + kProdThreads * j >= (uint32_t)kMmaRows) { c[2 * j] = 0; c[2 * j + 1] = 0; }
                    else if (r0 + 1 < A.row_end) {
                        const uint4 v = ldg_stream_v4(reinterpret_cast<const uint4 *>(A.codes + r0));
                        c[2 * j] = (uint64_t)v.y << 32 | v.x; c[2 * j + 1] = (uint64_t)v.w << 32 | v.z;
                    } else { c[2 * j] = r0 < A.row_end ? A.codes[r0] : 0; c[2 * j + 1] = 0; }
                }
            };
            uint64_t cur[2 * kProdRows], nxt[2 * kProdRows];
#pragma unroll
            for (int j = 0; j < 2 * kProdRows; ++j) nxt[j] = 0;
            if (blockIdx.x < n_tiles) load_tile(blockIdx.x, cur);
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t s = it % kMmaStages, ph = (it / kMmaStages) & 1;
                if (tile + gridDim.x < n_tiles) load_tile(tile + gridDim.x, nxt);   // next tile's codes fly while this one is expanded
                mbar_wait_sleep(&cempty[s], ph ^ 1);
#pragma unroll
                for (int j = 0; j < kProdRows; ++j)
                    if (kMmaRows % kProdThreads == 0 || t + kProdThreads * j < (uint32_t)kMmaRows)
                        mma_store_code_row(sC + s * kMmaCBytes, t + kProdThreads * j, cur[2 * j], cur[2 * j + 1]);
                fence_proxy_async_smem();
                mbar_arrive(&cfull[s]);
#pragma unroll
                for (int j = 0; j < 2 * kProdRows; ++j) cur[j] = nxt[j];
            }
        }
    } else {
        const uint32_t quad = warp & 3, part = (warp - 1 - kProdWarps) >> 2;
        if constexpr (kStagger) {
            // ===== epilogue, two groups out of phase: group g (8 warps) serves accumulator stage g, i.e. every second item =====
            // warp -> TMEM lane quadrant (warp % 4) and 128 of the stage's 256 columns, read as two 64-column strips into two
            // register images.  The stage goes back to the MMA issuer as soon as both strips have landed; the min/max work of this
            // group then overlaps the other group's barrier round trip and tcgen05.ld, which all sixteen warps of the lock-step
            // schedule (kStagger = false below) sit out together.  Needs 64 payload registers per thread: only the stage-image form
            // of the kernel (18 warps, 112 registers) has them.
            const uint32_t grp = part & 1, colh = part >> 1;
            const uint32_t taddr0 = tmem_base + ((quad * 32u) << 16) + grp * kMmaRows + colh * (2 * kMmaColsPerWarp);
            const uint32_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
            const uint2 *my_bnd = s_bnd + quad * 32 + lane;
            // the strip's first row, computed only when a strip fires: tile ti of this CTA, query tile mt
            auto strip_row = [&](uint32_t ti, uint32_t strip) {
                return A.row0 + (uint64_t)(blockIdx.x + ti * gridDim.x) * kMmaTileCodes + 2 * (2 * colh + strip) * kMmaColsPerWarp;
            };
            uint32_t aph = 0, ti = 0, mt = grp;   // items in issue order are (tile, query tile); this group takes every second one
            for (;;) {
                while (mt >= q_tiles) { mt -= q_tiles; ++ti; }
                if (ti >= my_tiles) break;
                const uint2 bnd = my_bnd[mt * kMmaQTile];
                mbar_wait_sleep(&tfull[grp], aph);
                tcgen05_fence_after();
                uint32_t p0[32], p1[32];   // register c = (D of column 2c+1) << 16 | (D of column 2c) & 0xFFFF
                if (!UCFP_DIAG_NO_LD(A)) {
                    tmem_ld64_pack16_async(taddr0, p0);
                    tmem_ld64_pack16_async(taddr0 + kMmaColsPerWarp, p1);
                    tmem_ld_wait(p0);
                    tmem_ld_tie(p1);
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[grp]);   // both strips have left TMEM: release the stage
                aph ^= 1;
                if (!UCFP_DIAG_NO_TEST(A) && hamming_mma_strip_test<!kPreExpanded>(p0, bnd.x, bnd.y))
                    hamming_mma_fire(p0, strip_row(ti, 0), mt * kMmaQTile + quad * 32 + lane, A, s_spill_n, s_q, s_kid);
                if (!UCFP_DIAG_NO_TEST(A) && hamming_mma_strip_test<!kPreExpanded>(p1, bnd.x, bnd.y))
                    hamming_mma_fire(p1, strip_row(ti, 1), mt * kMmaQTile + quad * 32 + lane, A, s_spill_n, s_q, s_kid);
                mt += 2;
            }
        } else {
        // ===== epilogue: warp -> TMEM lane quadrant (warp % 4) and 64 of the 256 columns =====
        uint32_t as = 0, aph = 0;   // accumulator stage in use and its mbarrier phase
        const uint32_t taddr0 = tmem_base + ((quad * 32u) << 16) + part * kMmaColsPerWarp;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint64_t first_row = A.row0 + (uint64_t)tile * kMmaTileCodes + 2 * part * kMmaColsPerWarp;
            for (uint32_t mt = 0; mt < q_tiles; ++mt) {
                const uint32_t q = mt * kMmaQTile + quad * 32 + lane;
                const uint2 bnd = s_bnd[q];
                mbar_wait_sleep(&tfull[as], aph);
                tcgen05_fence_after();
                const uint32_t taddr = taddr0 + as * kMmaRows;
                uint32_t p[32];   // register c = (D of column 2c+1) << 16 | (D of column 2c) & 0xFFFF
                if (!UCFP_DIAG_NO_LD(A)) { tmem_ld64_pack16_async(taddr, p); tmem_ld_wait(p); }
                const bool fired = !UCFP_DIAG_NO_TEST(A) && hamming_mma_strip_test<false>(p, bnd.x, bnd.y);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[as]);   // the accumulators now live in registers: release the stage first
                as ^= 1; aph ^= as ^ 1;
                if (fired) hamming_mma_fire(p, first_row, q, A, s_spill_n, s_q, s_kid);
            }
        }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && A.spill) A.spill_count[blockIdx.x] = *s_spill_n;
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc_512(tmem_base);
    }
}

#ifdef UCFP_HAMMING_EXPERIMENTS
#include "hamming_experiments.cuh"   // round-2 epilogue schedules that were measured and not adopted
#endif

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

// Called before c->size grows: rows [first_row, first_row + n) were just written.  The pair row of an odd first_row is rebuilt.
int hamming_on_append(ucfp_lane *ctx, ucfp_corpus *c, uint64_t first_row, uint64_t n) {
    if (!c->ham_ops || n == 0) return UCFP_OK;
    const uint64_t lo = first_row & ~1ULL, hi = first_row + n;
    const uint64_t pairs = (hi - lo + 1) / 2;
    // `size` = rows that exist once this call is done: the appended range's end, or the corpus size when rows are rebuilt in place
    const uint64_t size_after = hi > c->size ? hi : c->size;
    hamming_ops_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, ctx->stream>>>(static_cast<const uint64_t *>(c->rows), lo, hi, size_after,
                                                                                  reinterpret_cast<uint4 *>(c->ham_ops));
    count_launch(ctx);
    return check_launch("hamming_ops");
}

// Once per context (ucfp_init): opt the kernels into their dynamic shared memory and measure the occupancies the launch
// geometry depends on -- per-device properties that used to be set on every scan call.
int hamming_device_init(ucfp_ctx *ctx) {
    UCFP_CUDA_TRY(cudaFuncSetAttribute(compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8192));
    int occ = 0;
    UCFP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hamming_scan_kernel<false>, kScanThreads, sizeof(QSlot) * kMaxQueriesPerPass));
    ctx->ham_scan_occ = occ < 1 ? 1 : occ;
    UCFP_CUDA_TRY((cudaFuncSetAttribute(hamming_mma_scan_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem)));
    UCFP_CUDA_TRY((cudaFuncSetAttribute(hamming_mma_scan_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem)));
    UCFP_CUDA_TRY((cudaFuncSetAttribute(hamming_mma_scan_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem)));
    UCFP_CUDA_TRY((cudaFuncSetAttribute(hamming_mma_scan_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem)));
#ifdef UCFP_HAMMING_EXPERIMENTS
    UCFP_CUDA_TRY(cudaFuncSetAttribute(hamming_mma_scan2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(hamming_mma_scan2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(hamming_mma_scan3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(hamming_mma_scan3_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(hamming_mma_scan4_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(hamming_mma_scan4_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMmaSmem));
#endif
    UCFP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, exact_select_kernel<HammingKey>, 256, 0));
    ctx->ham_exact_occ = occ < 1 ? 1 : (occ > 4 ? 4 : occ);
    return UCFP_OK;
}

int hamming_scan(ucfp_lane *ctx, ucfp_corpus *c, const uint64_t *q_dev, size_t nq, size_t k, uint64_t *ids_out_dev,
                 uint32_t *dist_out_dev, bool emit_rows) {
    cudaStream_t st = ctx->stream;
    const uint64_t N = c->size;
    UCFP_REQUIRE(k <= 2048, UCFP_E_UNSUPPORTED, "hamming scan supports k <= 2048 (got %zu)", k);
    UCFP_REQUIRE(N <= kRowMask, UCFP_E_UNSUPPORTED, "corpus too large");
    if (N == 0) {
        size_t tot = nq * k;
        fill_sentinel_u32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ids_out_dev, dist_out_dev, tot);
        count_launch(ctx);
        UCFP_TRY(check_launch("fill_sentinel"));
        if (!ctx->xch) return UCFP_OK;   // an empty shard of a group scan still takes part in the bound exchanges below
    }
    uint32_t cap = 4096;
    while (cap < 4 * k) cap <<= 1;
    const uint64_t *codes = static_cast<const uint64_t *>(c->rows);
    const uint64_t *ids = c->id_mode == 1 ? c->ids : nullptr;

    const int scan_occ = ctx->owner->ham_scan_occ;
#ifdef UCFP_HAMMING_EXPERIMENTS
    static const long env_wait = getenv("UCFP_HAMMING_WAIT") ? atol(getenv("UCFP_HAMMING_WAIT")) : 0;
    static const long env_epi_w = getenv("UCFP_HAMMING_EPI_WARPS") ? atol(getenv("UCFP_HAMMING_EPI_WARPS")) : 16;
    static const long env_mma_v = getenv("UCFP_HAMMING_MMA_V") ? atol(getenv("UCFP_HAMMING_MMA_V")) : 1;
#elif defined(UCFP_HAMMING_DIAG)
    static const long env_wait = getenv("UCFP_HAMMING_DIAG") ? atol(getenv("UCFP_HAMMING_DIAG")) : 0;
#else
    constexpr long env_wait = 0;
#endif
    static const bool env_no_mma = getenv("UCFP_HAMMING_NO_MMA") != nullptr;   // developer switches: POPC scan only /
    static const bool env_no_ops = getenv("UCFP_HAMMING_NO_OPS") != nullptr;     // expand codes in the kernel although operand rows exist
    static const bool env_no_spill = getenv("UCFP_HAMMING_NO_SPILL") != nullptr;   // developer switch: settle every fire inside the scan (round-1 behaviour)
    static const bool env_stagger_exp = getenv("UCFP_HAMMING_STAGGER_EXP") ? atol(getenv("UCFP_HAMMING_STAGGER_EXP")) != 0 : true;   // same for the expansion form
    static const bool env_stagger = getenv("UCFP_HAMMING_STAGGER") ? atol(getenv("UCFP_HAMMING_STAGGER")) != 0 : true;   // developer switch: 0 = lock-step epilogue in the image form too
    // Batches above this size expand the codes in the kernel even when stage images exist: with six or more query tiles per stage the
    // expansion (three producer warps) hides behind the MMAs, and its HBM traffic is the algorithmic 8 B per code -- over 1 B codes the
    // image form's extra 24 GB per batch cost the clock 2 % under sw_power_cap.  Both forms run the two-group epilogue.  Measured per
    // 250 M codes (scripts/r2/gpu_step34-36.sh): 640 queries 6.28 ms on images / 6.54 expanding, 768: 7.31 / 6.87, 896: 8.40 / 7.80,
    // 1 024: 9.4 / 8.82; 1 B codes x 1 024 queries: 37.97 ms lock-step -> 34.36 ms.
    static const long env_img_maxq = getenv("UCFP_HAMMING_IMG_MAXQ") ? atol(getenv("UCFP_HAMMING_IMG_MAXQ")) : 5 * kMmaQTile;

    for (size_t q0 = 0; q0 < nq; q0 += kMaxQueriesPerPass) {
        const uint32_t nqp = (uint32_t)((nq - q0 < kMaxQueriesPerPass) ? nq - q0 : kMaxQueriesPerPass);
        UCFP_TRY(ctx->qstate.reserve(sizeof(QSlot) * nqp + sizeof(uint64_t) * nqp));
        UCFP_TRY(ctx->cand.reserve(sizeof(uint64_t) * (size_t)cap * nqp));
        UCFP_TRY(ctx->cand_count.reserve(sizeof(uint32_t) * nqp));
        UCFP_TRY(ctx->flags.reserve(sizeof(uint32_t) * (2 * nqp + 1)));
        QSlot *slots = ctx->qstate.as<QSlot>();
        uint64_t *kth = reinterpret_cast<uint64_t *>(slots + nqp);
        uint64_t *cand = ctx->cand.as<uint64_t>();
        uint32_t *count = ctx->cand_count.as<uint32_t>();
        uint32_t *flags = ctx->flags.as<uint32_t>();
        uint64_t *ids_out = ids_out_dev + q0 * k;
        uint32_t *dist_out = dist_out_dev + q0 * k;

        hamming_init_kernel<<<(nqp + 255) / 256, 256, 0, st>>>(q_dev + q0, nqp, slots, kth, count, flags);
        if (N == 0) {   // group scan, empty shard: contribute the trivial bound (64, none) to every exchange
            count_launch(ctx);
            SelectState none{cand, count, &slots[0].thr, 4, kth, flags, cap, flags + nqp};
            while (ctx->xch->done < kBoundExchanges) UCFP_TRY(exchange_bounds(ctx, none, nqp));
            ctx->xch->done = 0;
            continue;
        }
        static const long env_seed = getenv("UCFP_HAMMING_SEED") ? atol(getenv("UCFP_HAMMING_SEED")) : 0;   // developer knobs
        static const long env_mma_rows = getenv("UCFP_HAMMING_MMA_MIN_ROWS") ? atol(getenv("UCFP_HAMMING_MMA_MIN_ROWS")) : 0;
        uint32_t seed_rows = env_seed > 0 ? (uint32_t)env_seed : kSeedRows;
        if (seed_rows < 2 * k) seed_rows = (uint32_t)(2 * k);   // the seed must fill a k-list with room to spare, or its bound admits everything
        seed_rows = (seed_rows + kMmaTileCodes - 1) / kMmaTileCodes * kMmaTileCodes;   // chunk starts = seed x growth^i stay aligned to the 512-code stage images
        const uint32_t seed = (uint32_t)(N < seed_rows ? N : seed_rows);
        const uint64_t mma_min_rows = env_mma_rows > 0 ? (uint64_t)env_mma_rows : kMmaMinChunkRows;
        hamming_seed_kernel<<<dim3((seed + 255) / 256, nqp), 256, 0, st>>>(codes, seed, slots, cand, count, cap);
        count_launch(ctx, 2);

        SelectState sel{cand, count, &slots[0].thr, 4, kth, flags, cap, flags + nqp, ctx->stats.as<unsigned long long>() + 1, emit_rows ? 1 : 0, /*small_keys=*/1};
        auto compact = [&](bool final_pass) {
            compact_lists(sel, nqp, (uint32_t)k, ids, c->id_base, final_pass, 0u, ids_out, dist_out, st);
            count_launch(ctx, 2);
        };
        compact(seed == N);

        // small batches are HBM-bound: fewer, larger chunks; large batches tighten thr more often
        const bool streaming = nqp <= 16;   // HBM-bound regime: fewer and larger chunks
        const bool use_mma = nqp >= kMmaMinQueries && !env_no_mma;
        static const long env_growth = getenv("UCFP_HAMMING_GROWTH") ? atol(getenv("UCFP_HAMMING_GROWTH")) : 0;
        // a chunk of g x (rows seen) admits about g x k rows per query; the lists hold cap entries including the k kept ones
        const uint64_t growth_cap = (cap - k) / (2 * k) > 2 ? (cap - k) / (2 * k) : 2;
        const uint64_t growth_want = streaming ? 64 : (env_growth > 1 ? (uint64_t)env_growth : 8);
        const uint64_t growth = growth_want < growth_cap ? growth_want : growth_cap;
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
            if (use_mma && n >= mma_min_rows) {
                const uint64_t tiles = (n + kMmaTileCodes - 1) / kMmaTileCodes;
                const unsigned mma_grid = (unsigned)(tiles < (uint64_t)ctx->sm_count ? tiles : (uint64_t)ctx->sm_count);
                ProfScope ps(ctx, UCFP_PROF_HAMMING_SCAN, 8.0 * (double)n * nqp);
                ProfScope pt(ctx, UCFP_PROF_HAMMING_TENSOR, 64.0 * (double)n * nqp);
                UCFP_TRY(ctx->spill.reserve((size_t)mma_grid * kSpillCap * sizeof(uint4) + 4 * (size_t)mma_grid + 256));
                uint4 *spill = ctx->spill.as<uint4>();
                uint32_t *spill_count = reinterpret_cast<uint32_t *>(spill + (size_t)mma_grid * kSpillCap);
                const MmaScanArgs margs{codes, reinterpret_cast<const uint4 *>(c->ham_ops), ids, c->id_base, pos, pos + n, slots, kth, nqp, cand, count, cap, (uint32_t)env_wait,
                                        env_no_spill ? nullptr : spill, spill_count, kSpillCap};
                // With all eight query tiles in use a 512-code stage lasts ~3 300 clk and the in-kernel expansion hides completely
                // behind it (measured 41.8 vs 43.0 ms per 1 B rows); below that the ready-made images win (7.6 vs 13.5 ms at 64-128 queries).
                const bool have_images = c->ham_ops && !env_no_ops && pos % kMmaTileCodes == 0;
                bool launched = false;
#ifdef UCFP_HAMMING_EXPERIMENTS
                launched = true;
                if (have_images && env_mma_v == 4 && env_epi_w == 8) hamming_mma_scan4_kernel<8><<<mma_grid, 32 * 10, kMmaSmem, st>>>(margs);
                else if (have_images && env_mma_v == 4) hamming_mma_scan4_kernel<16><<<mma_grid, 32 * 18, kMmaSmem, st>>>(margs);
                else if (have_images && env_mma_v == 3 && env_epi_w == 8) hamming_mma_scan3_kernel<8><<<mma_grid, 32 * 10, kMmaSmem, st>>>(margs);
                else if (have_images && env_mma_v == 3) hamming_mma_scan3_kernel<16><<<mma_grid, 32 * 18, kMmaSmem, st>>>(margs);
                else if (have_images && env_mma_v == 2 && env_epi_w == 8) hamming_mma_scan2_kernel<8><<<mma_grid, 32 * 10, kMmaSmem, st>>>(margs);
                else if (have_images && env_mma_v == 2) hamming_mma_scan2_kernel<16><<<mma_grid, 32 * 18, kMmaSmem, st>>>(margs);
                else launched = false;
#endif
                if (launched) {}
                else if (have_images && (long)nqp <= env_img_maxq) {
                    if (env_stagger) hamming_mma_scan_kernel<true, true><<<mma_grid, kMmaStaggerThreads, kMmaSmem, st>>>(margs);
                    else hamming_mma_scan_kernel<true, false><<<mma_grid, kMmaThreads, kMmaSmem, st>>>(margs);
                } else if (env_stagger_exp && nqp > 5 * kMmaQTile) {   // three producer warps keep up from six query tiles per stage on (measured: 768 queries 7.57 -> 6.95 ms per 250 M codes, 512 queries 5.35 -> 5.84)
                    hamming_mma_scan_kernel<false, true><<<mma_grid, kMmaStaggerExpThreads, kMmaSmem, st>>>(margs);
                } else hamming_mma_scan_kernel<false, false><<<mma_grid, kMmaThreads, kMmaSmem, st>>>(margs);
                if (!launched && !env_no_spill) {   // the strips this launch parked are settled before the compaction
                    recheck_parked_kernel<<<dim3(mma_grid, kRecheckSlices), 256, 0, st>>>(spill, spill_count, kSpillCap, codes, ids, c->id_base, pos + n, slots, kth, cand, count, cap);
                    count_launch(ctx);
                }
            } else {
                ProfScope ps(ctx, UCFP_PROF_HAMMING_SCAN, 8.0 * (double)n * nqp);
                hamming_scan_kernel<false><<<(unsigned)grid, kScanThreads, sizeof(QSlot) * nqp, st>>>(
                    codes, ids, c->id_base, pos, n, slots, kth, nqp, (uint32_t)q_groups, cand, count, cap, nullptr);
            }
            count_launch(ctx);
            pos += n;
            compact(pos == N);
            // group scan: fold in the other shards' bounds once this shard has seen enough rows for the next exchange
            if (ctx->xch && pos < N)
                while (ctx->xch->done < kBoundExchanges && pos >= kBoundExchangeRows[ctx->xch->done]) UCFP_TRY(exchange_bounds(ctx, sel, nqp));
            chunk = chunk * growth < max_chunk ? chunk * growth : max_chunk;
        }
        if (ctx->xch) {   // exchanges this shard was too short for: the other ranks are waiting in theirs
            while (ctx->xch->done < kBoundExchanges) UCFP_TRY(exchange_bounds(ctx, sel, nqp));
            ctx->xch->done = 0;
        }
        UCFP_TRY(check_launch("hamming scan"));
        // Queries whose candidate list overflowed somewhere (device-side decisions, no host sync): up to kRescanRounds streaming
        // passes over the corpus for the flagged queries alone, under the bound their truncated lists produced; the launches
        // return at once when nothing is flagged.  What is still flagged afterwards goes to the exact multi-pass selection.
        UCFP_TRY(stats_add_flags(ctx, flags, nqp));
        for (int round = 0; round < rescan_rounds(); ++round) {
            hamming_scan_kernel<true><<<(unsigned)(ctx->sm_count * scan_occ), kScanThreads, sizeof(QSlot) * nqp, st>>>(
                codes, ids, c->id_base, 0, N, slots, kth, nqp, 1u, cand, count, cap, flags);
            compact_rescanned(sel, nqp, (uint32_t)k, ids, c->id_base, 0u, ids_out, dist_out, st);
            count_launch(ctx, 2);
        }
        UCFP_TRY(exact_select_fallback(ctx, c, ctx->owner->ham_exact_occ, HammingKey{codes, slots, 0, 0}, flags, nqp, (uint32_t)k, 0u, ids_out, dist_out, emit_rows ? 1 : 0));
    }
    return UCFP_OK;
}

}  // namespace ucfp
