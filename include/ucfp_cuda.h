/*
 * ucfp_cuda.h -- C ABI of libucfp_cuda.so, the B200 (sm_100a) implementation of
 * UCFP's data-parallel fingerprint hot path.
 *
 * This is the boundary the reference's Rust host binds over FFI (crate
 * `ucfp-cuda`, see INTEGRATION.md and rust/ucfp-cuda/).  Plain pointers and
 * sizes only: no C++ types, no torch types, no ownership transfer.  Every
 * entry point returns 0 (UCFP_OK) or a negative UCFP_E_* code and never throws
 * or aborts (the reference builds with panic = "abort", Cargo.toml:180, so
 * nothing may unwind across this line).  The message for the last error on the
 * calling thread is available from ucfp_last_error().
 *
 * Buffers may live in host memory (pageable or pinned) or in device memory of
 * the context's GPU; the library inspects each pointer
 * (cudaPointerGetAttributes) and stages host buffers itself.
 *
 * Threading (the reference calls this path from a multi-thread tokio runtime
 * with <= 512 requests in flight, src/bin/ucfp.rs:207,262-267): every entry
 * point is thread-safe.  A call leases a "lane" of its context -- a private
 * CUDA stream plus private scratch, up to 16 per context -- so calls from
 * different host threads run concurrently on the GPU.  A corpus is read under
 * a shared lock (any number of concurrent scans) and mutated (append, clear,
 * delete, upsert, reserve, refresh) under an exclusive one.  In this default
 * ("pooled") mode every call returns with its work complete, whatever memory
 * its outputs live in.
 * ucfp_ctx_set_stream switches the context to "shared-stream" mode for hosts
 * that own a CUDA stream (a torch stream in the test harness): all calls are
 * then enqueued on that stream, one at a time, and calls whose outputs are all
 * device buffers return asynchronously; calls with host outputs synchronise.
 *
 * There is NO CPU fallback: without a usable sm_100 device every call fails
 * with UCFP_E_CUDA.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   hashing seam  src/modality/image.rs:68-70   ImageFingerprinter::fingerprint_with_preprocess
 *                 src/modality/image.rs:175-179 FingerprinterContext::fingerprint_with_algorithm_and_preprocess
 *   scan seam     src/index/mod.rs:29-35        IndexBackend::knn
 *                 src/index/embedded/mod.rs:268-360 EmbeddedBackend::knn (+ :454-495 helpers)
 *   new (absent from the reference): Hamming and MinHash-Jaccard top-k.
 */
#ifndef UCFP_CUDA_H
#define UCFP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define UCFP_API
#else
#define UCFP_API __attribute__((visibility("default")))
#endif

#define UCFP_ABI_VERSION 2

/* ---- status codes (map onto src/error.rs:9-61 on the Rust side) ---------- */
enum {
    UCFP_OK = 0,
    UCFP_E_INVALID = -1,     /* bad argument -> Error::Modality / Error::Index           */
    UCFP_E_CUDA = -2,        /* CUDA runtime / driver failure, or no sm_100 device       */
    UCFP_E_OOM = -3,         /* device or host allocation failed                         */
    UCFP_E_UNSUPPORTED = -4, /* valid request this build cannot serve -> Error::Unsupported */
    UCFP_E_STATE = -5,       /* object used in the wrong state (e.g. wrong corpus kind)  */
    UCFP_E_CAPACITY = -6     /* append beyond the corpus capacity                        */
};

/* id returned in unused result slots (fewer than k rows matched) */
#define UCFP_ID_NONE UINT64_MAX

typedef struct ucfp_ctx ucfp_ctx;       /* one per (process, GPU) */
typedef struct ucfp_corpus ucfp_corpus; /* one per (tenant, kind, dim): HBM-resident rows */

/* ---- context -------------------------------------------------------------- */

/* Binds a context to CUDA device `device`.  Fails with UCFP_E_CUDA when the
 * device is missing or is not compute capability 10.x. */
UCFP_API int ucfp_init(int device, ucfp_ctx **out);
UCFP_API void ucfp_destroy(ucfp_ctx *ctx);
/* Shared-stream mode: use the caller's cudaStream_t (passed as void*) for all subsequent work of this context, one
 * call at a time.  NULL is CUDA's default stream, exactly as in the runtime API. */
UCFP_API int ucfp_ctx_set_stream(ucfp_ctx *ctx, void *cuda_stream);
/* Back to pooled mode (the state after ucfp_init): per-call lanes with their own non-blocking streams. */
UCFP_API int ucfp_ctx_reset_stream(ucfp_ctx *ctx);
/* Blocks until everything enqueued by this context has finished. */
UCFP_API int ucfp_ctx_synchronize(ucfp_ctx *ctx);
UCFP_API int ucfp_abi_version(void);
/* Thread-local, never NULL, valid until the next failing call on this thread. */
UCFP_API const char *ucfp_last_error(void);
/* Number of kernels this library has launched on `ctx` since creation (bench.py's gpu_launches). */
UCFP_API uint64_t ucfp_ctx_kernel_launches(const ucfp_ctx *ctx);

/* Per-kernel timing for roofline reports.  Between _begin and _end the library brackets every launch of
 * its dominant kernels with CUDA events on the context's stream.  _end synchronises and returns, for
 * one kernel class, the summed device time, the summed ALGORITHMIC bytes (Hamming: 8 B x rows x queries
 * of each launch; Jaccard: 1024 B x rows x queries; image: 3*w*h + 408 B per image) or flops (cosine:
 * 2 x rows x dim x queries) and the number of launches.  UCFP_PROF_HAMMING_SCAN covers every scan launch of a
 * Hamming call; UCFP_PROF_HAMMING_TENSOR only the launches of the tensor-core scan (batches >= 64 queries), with
 * units = int8 operations the tensor pipe executes: 64 per (query, code) pair (one 64-element +-1 dot product per
 * TWO pairs, two codes being packed into one operand row).  _read returns the sums without ending the region. */
enum { UCFP_PROF_HAMMING_SCAN = 1, UCFP_PROF_JACCARD_SCAN = 2, UCFP_PROF_COSINE_SCAN = 3, UCFP_PROF_IMAGE_HASH = 4,
       UCFP_PROF_HAMMING_TENSOR = 5 };
UCFP_API int ucfp_ctx_profile_begin(ucfp_ctx *ctx);
UCFP_API int ucfp_ctx_profile_end(ucfp_ctx *ctx, int kernel_class, double *kernel_ms, double *alg_units,
                                  uint64_t *launches);
UCFP_API int ucfp_ctx_profile_read(ucfp_ctx *ctx, int kernel_class, double *kernel_ms, double *alg_units,
                                   uint64_t *launches);

/* ---- image hashing seam ---------------------------------------------------
 * Replaces the calls into imgfprint at src/modality/image.rs:68-70 and
 * :175-179 AFTER host-side decode.  The host keeps: decoding, EXIF orientation,
 * the size/dimension guards of build_image_preprocess
 * (src/server/handlers.rs:307-319), BLAKE3 `exact`, and assembling the
 * 168-byte ImageFingerprint / 536-byte MultiHashFingerprint structs. */

enum { UCFP_ALGO_AHASH = 1u, UCFP_ALGO_PHASH = 2u, UCFP_ALGO_DHASH = 4u, UCFP_ALGO_MULTI = 7u };

typedef struct ucfp_image_desc {
    const uint8_t *pixels; /* interleaved RGB8, host or device memory      */
    uint32_t width;        /* pixels, 4 <= width                            */
    uint32_t height;       /* pixels, 4 <= height                           */
    uint64_t stride;       /* bytes between rows, >= 3 * width              */
} ucfp_image_desc;

/* One algorithm's hashes for one image: the `global_hash` and `block_hashes`
 * fields of imgfprint's ImageFingerprint (layout documented in
 * web/src/lib/components/charts/ImageHashView.svelte:2-5).  Block 4r+c covers
 * rows [r*h/4,(r+1)*h/4) x columns [c*w/4,(c+1)*w/4). */
typedef struct ucfp_hash17 {
    uint64_t global_hash;
    uint64_t block_hashes[16];
} ucfp_hash17;

/* The three algorithms in MultiHashFingerprint order (ahash, phash, dhash;
 * web/src/lib/components/charts/AlgorithmView.svelte:30-37).  Algorithms not
 * selected by algo_mask are left zero. */
typedef struct ucfp_image_hashes {
    ucfp_hash17 ahash;
    ucfp_hash17 phash;
    ucfp_hash17 dhash;
} ucfp_image_hashes; /* 408 bytes */

/* Hash `n` decoded images.  `out` (n entries) and `status` (n entries, may be
 * NULL) may be host or device memory.  A bad image (NULL pixels, dimension
 * below 4, stride too small, or a shape the kernels cannot stage) sets its
 * status to a UCFP_E_* code and zeroes its output without failing the batch;
 * the call's own return value reports only batch-level failures. */
UCFP_API int ucfp_image_hash_batch(ucfp_ctx *ctx, const ucfp_image_desc *imgs, size_t n, uint32_t algo_mask,
                                   ucfp_image_hashes *out, int32_t *status);

/* Same, for `n` equally-sized images laid out `image_stride` bytes apart starting
 * at `pixels` (the batch-ingest layout; one descriptor for the whole batch).  The
 * buffer must lie in ONE kind of memory, all host or all device: its first and last
 * byte are checked once, not every image (UCFP_E_INVALID when they disagree). */
UCFP_API int ucfp_image_hash_uniform(ucfp_ctx *ctx, const uint8_t *pixels, size_t n, uint32_t width, uint32_t height,
                                     uint64_t row_stride, uint64_t image_stride, uint32_t algo_mask,
                                     ucfp_image_hashes *out);

/* Batch ingest from ENCODED bytes.  The n JPEG bitstreams (host memory) are decoded on the device by
 * nvJPEG and hashed where they land: decoded pixels never cross PCIe.  Replaces, for JPEG uploads, the host decode in front
 * of src/modality/image.rs:68-70 (the reference's benchmark includes that decode, benches/end_to_end.rs:40-53).
 * status[i] (host memory, required): UCFP_OK; UCFP_E_UNSUPPORTED = not a JPEG or one nvJPEG refuses (decode it on the host
 * and use ucfp_image_hash_batch); UCFP_E_INVALID = corrupt bitstream / smaller than 4 px.  EXIF orientation is NOT applied
 * (hosts that honour it rotate the 8x8 / 9x8 / 32x32 semantics themselves or decode such files on the host).
 * dims_out (may be NULL): 2 n u32 {width, height}.  pixels_out (may be NULL; host or device, pixels_capacity bytes): the
 * decoded RGB8 of the successfully decoded images, tightly packed, image after image.  nvJPEG is loaded on first use;
 * UCFP_E_UNSUPPORTED when it is not installed. */
UCFP_API int ucfp_image_hash_jpeg_batch(ucfp_ctx *ctx, const uint8_t *const *jpegs, const size_t *lengths, size_t n, uint32_t algo_mask,
                                        ucfp_image_hashes *out, int32_t *status, uint32_t *dims_out, uint8_t *pixels_out,
                                        size_t pixels_capacity);

/* ---- corpus ---------------------------------------------------------------- */

enum {
    UCFP_KIND_HAMMING64 = 1,  /* row = one u64 code (global_hash @32 of ImageFingerprint) */
    UCFP_KIND_MINHASH128 = 2, /* row = 128 u64 slots (payload @8 of txtfp MinHashSig<128>, src/modality/text.rs:200-204) */
    UCFP_KIND_COSINE = 3,     /* row = dim f32 (Record::embedding, src/core/mod.rs:58)     */
    UCFP_KIND_MULTIHASH = 4   /* row = ucfp_image_hashes: the 51 hash words of a `multi` bundle (ahash | phash | dhash, each global + 16
                                 blocks), for the block-hash-aware re-rank (ucfp_scan_multihash).  ucfp_corpus_append_strided on this kind
                                 takes 536-byte MultiHashFingerprint records (field_offset = offset of the record's `exact`) and gathers the
                                 three 136-byte hash runs at +64, +232, +400.  HBM per row of capacity: 408 + 8 + 32 (PHash side corpus) */
};

/* Allocates HBM for up to `capacity` rows.  `dim` is used by UCFP_KIND_COSINE only.  Bytes per row of capacity, side
 * arrays included: HAMMING64 8 + 32 (operand rows of the tensor-core scan; optional -- if that allocation is refused the
 * corpus still works and the scan expands codes on the fly), +8 once explicit ids are appended; MINHASH128 1024 + 256
 * (two sketch planes); COSINE 4*dim + 2*dim_pad + 4. */
UCFP_API int ucfp_corpus_create(ucfp_ctx *ctx, int kind, uint32_t dim, uint64_t capacity, ucfp_corpus **out);
UCFP_API void ucfp_corpus_destroy(ucfp_corpus *c);
/* Appends n rows (copied).  ids == NULL means record_id = id_base + row index, where id_base is the
 * value set by ucfp_corpus_set_id_base (default 0) -- the layout a range-sharded index uses; a corpus is
 * either all-explicit or all-implicit.  UCFP_ID_NONE is not a valid record id. */
UCFP_API int ucfp_corpus_append(ucfp_corpus *c, const uint64_t *ids, const void *rows, uint64_t n);
/* Bulk hydration from stored records: appends n rows, row i being the row-sized field at byte
 * `field_offset` of the record at `records + i * record_stride` -- e.g. the u64 global_hash at offset 32 of
 * 168-byte ImageFingerprints, the PHash global hash at offset 232 of 536-byte multi bundles, or the 1024-byte
 * slot payload at offset 8 of 1032-byte MinHashSig<128> blobs laid out back to back as read from the
 * fingerprints table.  `records` may be host or device memory; ids as in ucfp_corpus_append. */
UCFP_API int ucfp_corpus_append_strided(ucfp_corpus *c, const uint64_t *ids, const void *records, uint64_t record_stride,
                                        uint64_t field_offset, uint64_t n);
UCFP_API int ucfp_corpus_set_id_base(ucfp_corpus *c, uint64_t id_base);
/* Insert-or-replace by record id -- IndexBackend::upsert, src/index/mod.rs:20-22 ("Insert-or-replace by (tenant_id,
 * record_id)"): rows whose id is already resident are overwritten in place (side arrays re-derived), the others are
 * appended; within one batch the last occurrence of an id wins.  A full corpus grows by itself (reallocation + device-to-
 * device copy, at least doubling; UCFP_E_OOM when HBM refuses).  ids: host or device memory; rows: host or device.  An
 * implicit-id corpus becomes an explicit-id corpus (ids = id_base + row materialised) the first time it is mutated
 * through upsert or delete.  *n_replaced (may be NULL) = rows overwritten. */
UCFP_API int ucfp_corpus_upsert(ucfp_corpus *c, const uint64_t *ids, const void *rows, uint64_t n, uint64_t *n_replaced);
/* Removes the rows with these record ids -- IndexBackend::delete, src/index/mod.rs:23-25 (idempotent: unknown ids are
 * ignored).  The rows are found on the device (one pass over the id column), the corpus's LAST rows move into the freed
 * slots and their side arrays are re-derived: the corpus stays dense, no scan ever sees a tombstone, and since every
 * scan's total order breaks ties by record id the results do not depend on the row order.  *n_removed may be NULL. */
UCFP_API int ucfp_corpus_delete(ucfp_corpus *c, const uint64_t *ids, uint64_t n, uint64_t *n_removed);
/* Grows the allocation to at least `capacity` rows (no-op when it already is that large). */
UCFP_API int ucfp_corpus_reserve(ucfp_corpus *c, uint64_t capacity);
UCFP_API uint64_t ucfp_corpus_capacity(const ucfp_corpus *c);
UCFP_API int ucfp_corpus_clear(ucfp_corpus *c);
UCFP_API uint64_t ucfp_corpus_size(const ucfp_corpus *c);
/* Bench/test support: appends n synthetic rows generated on the device with the counter PRNG of
 * docs/HASH_SPEC.md section 8 (row r, word j = splitmix64(seed, (start_row + r) * words_per_row + j)),
 * implicit ids.  HAMMING64 and MINHASH128 only. */
UCFP_API int ucfp_corpus_append_synthetic(ucfp_corpus *c, uint64_t seed, uint64_t start_row, uint64_t n);
/* Device pointer to the resident rows (view for tests/bench planting), or NULL. */
UCFP_API void *ucfp_corpus_device_rows(ucfp_corpus *c);
/* Re-derives the side arrays (Hamming tensor-scan operand rows; MinHash sketches; cosine norms and bf16 copies) of
 * all resident rows.  Call it
 * after rows were modified in place through ucfp_corpus_device_rows. */
UCFP_API int ucfp_corpus_refresh(ucfp_corpus *c);

/* ---- scans ------------------------------------------------------------------
 * All results are ordered best first with the total order stated; slots beyond
 * the number of matching rows hold UCFP_ID_NONE and the sentinel given.  nq == 0
 * or k == 0 is a successful no-op, as in EmbeddedBackend::knn
 * (src/index/embedded/mod.rs:275). */

/* dist = popcount(q ^ code); order (dist asc, record_id asc); sentinel dist = UINT32_MAX. */
UCFP_API int ucfp_scan_hamming(ucfp_corpus *c, const uint64_t *queries, size_t nq, size_t k,
                               uint64_t *ids_out, uint32_t *dist_out);
/* matches = #{i < 128 : q[i] == row[i]}; order (matches desc, record_id asc); sentinel UINT32_MAX.
 * queries = nq x 128 u64. */
UCFP_API int ucfp_scan_jaccard(ucfp_corpus *c, const uint64_t *queries, size_t nq, size_t k,
                               uint64_t *ids_out, uint32_t *matches_out);
/* score = dot(q, v) / (|q| * |v|) in f32 as EmbeddedBackend::knn; order (score desc, record_id asc);
 * rows and queries with zero norm never match (:284, :328); sentinel score = -inf.  queries = nq x dim f32. */
UCFP_API int ucfp_scan_cosine(ucfp_corpus *c, const float *queries, size_t nq, size_t k,
                              uint64_t *ids_out, float *score_out);

/* ---- multi-hash compare and re-rank (docs/HASH_SPEC.md section 10) --------------------------------------------
 * The compare-time MultiHashConfig of the reference (src/modality/image.rs:21-24, 90-104; fields src/server/dto.rs:462-480;
 * defaults web/src/lib/docs/api-reference-image.md:51-62).  NULL = the defaults 0.1 / 0.4 / 0.3 / 0.1 / 0.1 / 12. */
typedef struct ucfp_multihash_config {
    float ahash_weight, phash_weight, dhash_weight; /* per-algorithm weights in [0, 1]; only ratios matter          */
    float global_weight, block_weight;              /* global-hash vs block-hash similarity inside every algorithm  */
    uint32_t block_distance_threshold;              /* a block matches when its Hamming distance is <= this (<= 64) */
} ucfp_multihash_config;
/* Re-rank: the k_prime rows nearest to each query's PHash global hash (Hamming scan, order of ucfp_scan_hamming) are scored
 * with the blended global + block similarity against the whole query bundle; the best k <= k_prime <= 2048 are returned
 * under (score desc, record_id asc), unused slots UCFP_ID_NONE / -inf.  queries: nq bundles, host or device. */
UCFP_API int ucfp_scan_multihash(ucfp_corpus *c, const ucfp_image_hashes *queries, size_t nq, size_t k_prime, size_t k,
                                 const ucfp_multihash_config *cfg, uint64_t *ids_out, float *score_out);
/* score_out[i] = blended similarity of bundles a[i] and b[i] (what imgfprint's compare would be asked; spec section 10). */
UCFP_API int ucfp_multihash_compare(ucfp_ctx *ctx, const ucfp_image_hashes *a, const ucfp_image_hashes *b, size_t n,
                                    const ucfp_multihash_config *cfg, float *score_out);

/* Limits: k <= 2048 (Hamming, Jaccard) and k <= 1024 (cosine) per call -- larger k returns UCFP_E_UNSUPPORTED; callers
 * that mirror IndexBackend::knn clamp k to min(k, ucfp_corpus_size) first (EmbeddedBackend::knn returns min(k, N) hits). */

/* ---- query batcher ----------------------------------------------------------
 * The reference serves one query per request with up to 512 requests in flight (src/bin/ucfp.rs:262-267,
 * handlers::query src/server/handlers.rs:143-187).  One query per scan call wastes the GPU: a scan of N rows costs the
 * same HBM pass for 1 query as for 64.  A batcher coalesces the single-query calls of many host threads into batched
 * scans: ucfp_batcher_query blocks its caller; worker threads owned by the batcher collect the queries that arrive while
 * the previous batch is scanning (at most max_batch; a first-in-line query waits at most max_delay_us for company) and
 * run them as ONE scan with k = the largest k of the batch; every caller gets the first k entries of its query's list
 * (the top-k' list is a prefix of the top-k list under the scans' total orders). */
typedef struct ucfp_batcher ucfp_batcher;
UCFP_API int ucfp_batcher_create(ucfp_corpus *c, uint32_t max_batch, uint32_t max_delay_us, ucfp_batcher **out);
UCFP_API void ucfp_batcher_destroy(ucfp_batcher *b);
/* query = one row of the corpus's kind (8 B code, 128 u64 slots, dim f32) in HOST memory; ids_out[k], keys_out[k] (u32
 * distances / matches, or f32 scores) in HOST memory.  Thread-safe; blocks until the result is there. */
UCFP_API int ucfp_batcher_query(ucfp_batcher *b, const void *query, size_t k, uint64_t *ids_out, void *keys_out);
UCFP_API int ucfp_batcher_stats(const ucfp_batcher *b, uint64_t *queries, uint64_t *batches, uint64_t *largest_batch);

/* ---- multi-GPU group --------------------------------------------------------
 * Record-range shards over the GPUs of one box: rank r holds rows [r*N/G, (r+1)*N/G) with GLOBAL record ids
 * (ucfp_corpus_set_id_base or explicit ids); every rank scans its shard against the whole query batch, the per-rank
 * top-k lists are exchanged as packed 16-byte (id, key) records in ONE NCCL all-gather and every rank runs the same
 * deterministic merge, so the result is byte-identical to the single-corpus scan.  Optionally (environment variable
 * UCFP_GROUP_EXCHANGES=1..4, read when the group is created, the same on every rank) the ranks also exchange their
 * per-query admission bounds at up to four chunk boundaries of a Hamming or Jaccard batch, so that every shard filters at
 * the best bound any rank has found so far; off by default: on 8 B200s it cost more than it saved.
 * NCCL is loaded at the first group call (dlopen of libnccl.so.2, the copy already in the process if there is one),
 * not at library load: hosts that never form a group do not need it. */
typedef struct ucfp_group ucfp_group;
/* Single process driving n GPUs (the FFI use-case): one context per device, ncclCommInitAll over them. */
UCFP_API int ucfp_group_create(const int *devices, int n, ucfp_group **out);
/* One process per GPU: rank 0 obtains an id, ships its 128 bytes to the other ranks by any means, every rank joins
 * with its own context. */
UCFP_API int ucfp_group_unique_id(void *id128);
UCFP_API int ucfp_group_join(ucfp_ctx *ctx, const void *id128, int rank, int world, ucfp_group **out);
UCFP_API void ucfp_group_destroy(ucfp_group *g);
UCFP_API int ucfp_group_local_size(const ucfp_group *g);           /* GPUs this process drives */
UCFP_API int ucfp_group_world_size(const ucfp_group *g);
UCFP_API ucfp_ctx *ucfp_group_ctx(ucfp_group *g, int local_rank);  /* owned by the group when created by ucfp_group_create */
/* corpora[i] = shard of local rank i (created on ucfp_group_ctx(g, i)).  queries: host memory, or device memory of any
 * local GPU.  Outputs: host memory, or device memory of one local GPU.  Collective over the WORLD: in the multi-process
 * form every process calls it with the same queries, nq and k. */
UCFP_API int ucfp_group_scan_hamming(ucfp_group *g, ucfp_corpus *const *corpora, const uint64_t *queries, size_t nq, size_t k,
                                     uint64_t *ids_out, uint32_t *dist_out);
UCFP_API int ucfp_group_scan_jaccard(ucfp_group *g, ucfp_corpus *const *corpora, const uint64_t *queries, size_t nq, size_t k,
                                     uint64_t *ids_out, uint32_t *matches_out);
UCFP_API int ucfp_group_scan_cosine(ucfp_group *g, ucfp_corpus *const *corpora, const float *queries, size_t nq, size_t k,
                                    uint64_t *ids_out, float *score_out);

/* Diagnostics of the most recent scan on this context (synchronises the stream): how many of its queries
 * overflowed their candidate list and were recomputed.  0 on the fast path.  A Hamming / Jaccard query that overflows
 * is first scanned again, together with the other flagged queries of its batch, under the bound its truncated lists
 * produced (at most two such rounds, one streaming pass over the corpus each); only a query that still overflows, and
 * every overflowing cosine query, goes to the exact multi-pass selection (~10 passes over the corpus per query). */
UCFP_API int ucfp_ctx_last_scan_fallbacks(ucfp_ctx *ctx, uint64_t *queries_recomputed);
/* Same, plus the longest candidate list any query of that scan (Hamming / Jaccard) accumulated between two compactions; the
 * lists hold 4096 entries (more for k > 1024), a longer one overflows.  A robustness gauge for skewed corpora: clustered
 * near-duplicates, floods of identical codes. */
UCFP_API int ucfp_ctx_last_scan_stats(ucfp_ctx *ctx, uint64_t *queries_recomputed, uint64_t *max_list_fill);
/* How many queries of the most recent scan were left to the exact multi-pass selection (a subset of the above). */
UCFP_API int ucfp_ctx_last_scan_exact_selects(ucfp_ctx *ctx, uint64_t *queries);

/* Merges `parts` per-shard result lists (each nq x k, best first, as written by a scan) into one
 * nq x k list under the same total order: the step after the NCCL all-gather of per-rank candidates.
 * ids_in / keys_in are laid out [part][query][k].  descending = 0 for Hamming distances, 1 for Jaccard
 * matches.  Every rank running this on the same gathered buffer gets byte-identical output. */
UCFP_API int ucfp_merge_topk_u32(ucfp_ctx *ctx, const uint64_t *ids_in, const uint32_t *keys_in, size_t parts,
                                 size_t nq, size_t k, int descending, uint64_t *ids_out, uint32_t *keys_out);
UCFP_API int ucfp_merge_topk_f32(ucfp_ctx *ctx, const uint64_t *ids_in, const float *scores_in, size_t parts,
                                 size_t nq, size_t k, uint64_t *ids_out, float *scores_out);

#ifdef __cplusplus
}
#endif
#endif /* UCFP_CUDA_H */
