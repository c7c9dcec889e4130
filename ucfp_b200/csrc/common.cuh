// common.cuh -- shared plumbing of libucfp_cuda.so: context/corpus objects, error
// reporting, host<->device staging.  No torch types; CUDA runtime only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <shared_mutex>
#include <string>
#include <vector>

#include "../../include/ucfp_cuda.h"

namespace ucfp {

// ---- error reporting -------------------------------------------------------
void set_error(const char *fmt, ...);

#define UCFP_CUDA_TRY(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::ucfp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return _e == cudaErrorMemoryAllocation ? UCFP_E_OOM : UCFP_E_CUDA;               \
        }                                                                                    \
    } while (0)

#define UCFP_TRY(expr)              \
    do {                            \
        int _rc = (expr);           \
        if (_rc != UCFP_OK) return _rc; \
    } while (0)

#define UCFP_REQUIRE(cond, code, ...)      \
    do {                                   \
        if (!(cond)) {                     \
            ::ucfp::set_error(__VA_ARGS__); \
            return (code);                 \
        }                                  \
    } while (0)

// ---- device scratch that grows on demand -----------------------------------
struct DevBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t want) {
        if (want <= bytes) return UCFP_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; bytes = 0;
        size_t rounded = (want + 255) & ~size_t(255);
        UCFP_CUDA_TRY(cudaMalloc(&ptr, rounded));
        bytes = rounded;
        return UCFP_OK;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

// A DevBuf that frees itself: for per-call temporaries that must not leak on an early error return.
struct ScopedDevBuf : DevBuf {
    ScopedDevBuf() = default;
    ScopedDevBuf(const ScopedDevBuf &) = delete;
    ScopedDevBuf &operator=(const ScopedDevBuf &) = delete;
    ~ScopedDevBuf() { release(); }
};

struct PinnedBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t want) {
        if (want <= bytes) return UCFP_OK;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr; bytes = 0;
        UCFP_CUDA_TRY(cudaMallocHost(&ptr, want));
        bytes = want;
        return UCFP_OK;
    }
    void release() { if (ptr) cudaFreeHost(ptr); ptr = nullptr; bytes = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

}  // namespace ucfp

// ---- the opaque objects of the C ABI ----------------------------------------
struct ucfp_ctx;

// Hook a group scan (group.cu) installs in the lane: between chunks the scan exchanges its per-query admission bounds with
// the other ranks through `allgather` (an NCCL all-gather on the lane's stream).
struct ucfp_exchange {
    void *comm = nullptr;
    int world = 1;
    int (*allgather)(void *comm, const void *send, void *recv, size_t bytes, cudaStream_t st) = nullptr;
    ucfp::DevBuf send, recv;
    int done = 0;   // exchanges performed in the query pass in hand
    int max_real = 0;   // of the kBoundExchanges slots of a pass, how many really exchange (the rest only count): see group.cu
};

// One stream plus the scratch a call needs.  An entry point LEASES a lane for its whole duration (api.cu, LaneLease), so
// nothing in here is shared between concurrent calls.  The kernels-side code takes the lane as `ctx` (it is "the context
// of this call": stream, scratch, SM count).
struct ucfp_lane {
    ucfp_ctx *owner = nullptr;
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr;   // created with the lane
    cudaStream_t stream = nullptr;       // the stream this lease runs on: own_stream, or the caller's (ucfp_ctx_set_stream)
    bool busy = false;                   // guarded by owner->mu
    // scratch of the scans
    ucfp::DevBuf q_dev, out_ids_dev, out_keys_dev, cand, cand_count, qstate, flags, misc;
    ucfp::DevBuf img_desc_dev, img_out_dev, img_status_dev, img_tables_dev, img_stage_dev;
    ucfp::DevBuf spill;                  // Hamming tensor scan: per-CTA queues of admitted pairs (64 KiB per CTA)
    ucfp::DevBuf mh_a, mh_b;             // multi-hash re-rank: query codes + coarse candidates, scored + merged lists
    ucfp::PinnedBuf pin_a, pin_b;
    ucfp::DevBuf stats;                  // u64[4] of this lane's last scan: [0] queries whose list overflowed (re-scanned), [1] longest list, [2] queries left to the exact selection
    ucfp_exchange *xch = nullptr;        // non-null during a group scan with more than one rank
};

constexpr int kUcfpMaxLanes = 16;

struct ucfp_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    // Lane pool.  Default ("pooled") mode: every call leases a free lane -- its own stream and scratch -- so calls from
    // different host threads run concurrently, and every call returns with its work complete.  After
    // ucfp_ctx_set_stream ("shared-stream" mode) all calls run on the caller's stream through lane 0, one at a time,
    // and calls whose outputs are all device buffers return asynchronously (the torch harness and bench.py use this).
    std::mutex mu;
    std::condition_variable cv;
    ucfp_lane *lanes[kUcfpMaxLanes] = {};
    int n_lanes = 0;
    bool shared_stream = false;
    cudaStream_t user_stream = nullptr;
    int last_scan_lane = 0;              // lane of the most recent scan (ucfp_ctx_last_scan_fallbacks)
    std::atomic<uint64_t> launches{0};   // kernels launched (gpu_launches in bench.py)
    // ucfp_ctx_profile_begin/_end
    std::atomic<bool> profiling{false};
    struct ProfRec { cudaEvent_t a, b; double units; int kind; };
    std::mutex prof_mu;
    std::vector<ProfRec> prof;
    // per-shape tap tables of image.cu (owned by it; freed by image_cache_destroy), shared by all lanes
    std::mutex image_mu;
    void *image_cache = nullptr;
    // nvJPEG decoder of ucfp_image_hash_jpeg_batch (owned by jpeg.cu; one decode batch at a time)
    std::mutex jpeg_mu;
    void *jpeg = nullptr;
    // launch parameters that depend on the device only, computed once by the *_device_init functions at ucfp_init
    int ham_scan_occ = 1, jac_scan_occ = 1, ham_exact_occ = 1, jac_exact_occ = 1, cos_exact_occ = 1;
};

struct ucfp_corpus {
    ucfp_ctx *ctx = nullptr;
    // Scans hold this shared for the whole call, append / clear / delete / upsert / refresh hold it exclusively.
    std::shared_mutex rw;
    int kind = 0;
    uint32_t dim = 0;
    uint64_t capacity = 0;
    uint64_t size = 0;
    uint64_t id_base = 0;
    int id_mode = 0;               // 0 undecided, 1 explicit ids, 2 implicit (id_base + row)
    void *rows = nullptr;          // HAMMING64: u64[cap]; MINHASH128: u64[cap][128]; COSINE: f32[cap][dim]
    uint64_t *ids = nullptr;       // u64[cap] when id_mode == 1
    // kind-specific side arrays
    uint8_t *ham_ops = nullptr;    // HAMMING64: s8[cap / 2][64], tensor-scan operand rows (codes 2r, 2r+1 as -a_k + 64 b_k); may be null
    uint8_t *mh_sketch = nullptr;  // MINHASH128: 2 planes of u8[cap][128]: byte 0 of every slot (scan prefilter), byte 1 (second check of survivors)
    void *cos_bf16 = nullptr;      // COSINE: bf16[cap][dim_pad] rows scaled to unit norm, for the tensor-core pass
    float *cos_inv_norm = nullptr; // COSINE: 1/|v| in f32 (0 for zero rows)
    uint32_t dim_pad = 0;
    ucfp_corpus *coarse = nullptr; // MULTIHASH: HAMMING64 side corpus of the PHash global hashes (word 17), same row order, implicit ids = rows
};

namespace ucfp {

enum class Mem { Host, Device };

// Classifies a user pointer.  Device = memory of ANY CUDA device or managed memory.
inline Mem classify(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return Mem::Host; }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? Mem::Device : Mem::Host;
}

inline void count_launch(ucfp_lane *ctx, uint64_t n = 1) { ctx->owner->launches.fetch_add(n, std::memory_order_relaxed); }

// Brackets one launch of a dominant kernel with events when profiling is on (no-ops otherwise).
struct ProfScope {
    ucfp_lane *ctx; cudaEvent_t a = nullptr, b = nullptr; double units; int kind;
    ProfScope(ucfp_lane *c, int kind_, double units_) : ctx(c), units(units_), kind(kind_) {
        if (!ctx->owner->profiling.load(std::memory_order_relaxed)) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; cudaGetLastError(); return; }
        cudaEventRecord(a, ctx->stream);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEventRecord(b, ctx->stream);
        try {
            std::lock_guard<std::mutex> lk(ctx->owner->prof_mu);
            ctx->owner->prof.push_back(ucfp_ctx::ProfRec{a, b, units, kind});
        } catch (...) { cudaEventDestroy(a); cudaEventDestroy(b); }
    }
};

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch %s failed: %s", what, cudaGetErrorString(e)); return UCFP_E_CUDA; }
    return UCFP_OK;
}

// ---- kernels-side entry points implemented in the per-path .cu files -------
// `ctx` is the lane the call has leased: its stream, its scratch.
int hamming_scan(ucfp_lane *ctx, ucfp_corpus *c, const uint64_t *q_dev, size_t nq, size_t k, uint64_t *ids_out_dev, uint32_t *dist_out_dev,
                 bool emit_rows = false);   // emit_rows: report row indices (ties still by record id)
int jaccard_scan(ucfp_lane *ctx, ucfp_corpus *c, const uint64_t *q_dev, size_t nq, size_t k, uint64_t *ids_out_dev, uint32_t *m_out_dev);
int cosine_scan(ucfp_lane *ctx, ucfp_corpus *c, const float *q_dev, size_t nq, size_t k, uint64_t *ids_out_dev, float *score_out_dev);
int hamming_on_append(ucfp_lane *ctx, ucfp_corpus *c, uint64_t first_row, uint64_t n);
int jaccard_on_append(ucfp_lane *ctx, ucfp_corpus *c, uint64_t first_row, uint64_t n);
int cosine_on_append(ucfp_lane *ctx, ucfp_corpus *c, uint64_t first_row, uint64_t n);
int multihash_on_append(ucfp_lane *ctx, ucfp_corpus *c, uint64_t first_row, uint64_t n);
int merge_u32(ucfp_lane *ctx, const uint64_t *ids_in, const uint32_t *keys_in, size_t parts, size_t nq, size_t k,
              int descending, uint64_t *ids_out, uint32_t *keys_out);
int merge_f32(ucfp_lane *ctx, const uint64_t *ids_in, const float *keys_in, size_t parts, size_t nq, size_t k,
              uint64_t *ids_out, float *keys_out);
int pack_topk(ucfp_lane *ctx, const uint64_t *ids, const void *keys32, size_t n, void *records_out);
int merge_packed(ucfp_lane *ctx, const void *records, size_t parts, size_t nq, size_t k, int key_is_f32, int descending, uint64_t *ids_out, void *keys_out);
int synth_fill_u64(ucfp_lane *ctx, uint64_t *dst_dev, uint64_t nwords, uint64_t seed, uint64_t start_word);
// pixels_mem: -1 = ask the driver where each image lives; 0 / 1 = every image is in host / device memory (the caller knows)
int image_hash_batch(ucfp_lane *ctx, const ucfp_image_desc *descs_host, size_t n, uint32_t algo_mask,
                     ucfp_image_hashes *out_dev, int32_t *status_host, int pixels_mem = -1);
// once per context, with the context's device current: kernel attributes (dynamic shared memory opt-in) and occupancies
int hamming_device_init(ucfp_ctx *ctx);
int jaccard_device_init(ucfp_ctx *ctx);
int cosine_device_init(ucfp_ctx *ctx);
int image_device_init(ucfp_ctx *ctx);
int merge_device_init(ucfp_ctx *ctx);

void image_cache_destroy(ucfp_ctx *ctx);
void jpeg_destroy(ucfp_ctx *ctx);
int stats_reset(ucfp_lane *ctx);
int stats_add_flags(ucfp_lane *ctx, const uint32_t *flags_dev, uint32_t nq);

// splitmix64 counter PRNG of docs/HASH_SPEC.md section 8
__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t splitmix64(uint64_t seed, uint64_t index) {
    return mix64(seed ^ ((index + 1) * 0x9E3779B97F4A7C15ULL));
}

}  // namespace ucfp
