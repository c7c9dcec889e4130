// api.cu -- the extern "C" surface of libucfp_cuda.so (declared in include/ucfp_cuda.h):
// context and lane pool, corpus lifetime, host/device staging, dispatch into the per-path kernels.
//
// Threading model (SURVEY 8b "Threading"): every entry point leases a LANE (one stream + its own scratch) for its
// duration, so calls from different host threads run concurrently; a corpus is read under a shared lock and
// mutated (append / clear / delete / upsert / refresh) under an exclusive one.  Nothing may unwind across the ABI:
// every body sits inside UCFP_API_BEGIN / UCFP_API_END, which turn any C++ exception into a status code.
#include <stdarg.h>

#include "api_util.cuh"

namespace ucfp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- lane pool ----------------------------------------------------------------------------------------------------
static ucfp_lane *lane_create(ucfp_ctx *ctx) {
    ucfp_lane *ln = new (std::nothrow) ucfp_lane();
    if (!ln) return nullptr;
    ln->owner = ctx; ln->device = ctx->device; ln->sm_count = ctx->sm_count; ln->smem_optin = ctx->smem_optin;
    if (cudaStreamCreateWithFlags(&ln->own_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete ln; return nullptr; }
    ln->stream = ln->own_stream;
    return ln;
}

static void lane_destroy(ucfp_lane *ln) {
    if (!ln) return;
    cudaStreamSynchronize(ln->own_stream);
    DevBuf *bufs[] = {&ln->q_dev, &ln->out_ids_dev, &ln->out_keys_dev, &ln->cand, &ln->cand_count, &ln->qstate, &ln->flags, &ln->misc,
                      &ln->img_desc_dev, &ln->img_out_dev, &ln->img_status_dev, &ln->img_tables_dev, &ln->img_stage_dev, &ln->stats, &ln->mh_a, &ln->mh_b, &ln->spill};
    for (DevBuf *b : bufs) b->release();
    ln->pin_a.release(); ln->pin_b.release();
    cudaStreamDestroy(ln->own_stream);
    delete ln;
}

LaneLease::LaneLease(ucfp_ctx *c) : ctx(c) {
    if (cudaGetDevice(&prev_device) != cudaSuccess) { prev_device = -1; cudaGetLastError(); }
    if (prev_device != ctx->device && cudaSetDevice(ctx->device) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaSetDevice(%d) failed", ctx->device);
        return;
    }
    std::unique_lock<std::mutex> lk(ctx->mu);
    for (;;) {
        if (ctx->shared_stream) {   // one call at a time, on the caller's stream, through lane 0
            if (!ctx->lanes[0]->busy) { lane = ctx->lanes[0]; lane->stream = ctx->user_stream; break; }
        } else {
            for (int i = 0; i < ctx->n_lanes && !lane; ++i)
                if (!ctx->lanes[i]->busy) lane = ctx->lanes[i];
            if (!lane && ctx->n_lanes < kUcfpMaxLanes) {
                ucfp_lane *fresh = lane_create(ctx);
                if (fresh) { ctx->lanes[ctx->n_lanes++] = fresh; lane = fresh; }
                else if (ctx->n_lanes == 0) { set_error("cannot create a stream"); return; }
            }
            if (lane) { lane->stream = lane->own_stream; break; }
        }
        ctx->cv.wait(lk);
    }
    lane->busy = true;
}

LaneLease::~LaneLease() {
    if (lane) {
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            lane->busy = false;
        }
        ctx->cv.notify_one();
    }
    if (prev_device >= 0 && prev_device != ctx->device) cudaSetDevice(prev_device);
}

// Makes `user` (host or device) readable on the device.  Host data is copied into `buf`.
int stage_in(ucfp_lane *ln, DevBuf &buf, const void *user, size_t bytes, const void **dev) {
    if (bytes == 0) { *dev = buf.ptr; return UCFP_OK; }
    if (classify(user) == Mem::Device) { *dev = user; return UCFP_OK; }
    UCFP_TRY(buf.reserve(bytes));
    UCFP_CUDA_TRY(cudaMemcpyAsync(buf.ptr, user, bytes, cudaMemcpyHostToDevice, ln->stream));
    *dev = buf.ptr;
    return UCFP_OK;
}

// Chooses where a kernel writes: straight into a device `user` buffer, or into `buf` for a host one.
int stage_out(DevBuf &buf, void *user, size_t bytes, void **dev, bool *is_host) {
    *is_host = classify(user) != Mem::Device;
    if (!*is_host) { *dev = user; return UCFP_OK; }
    UCFP_TRY(buf.reserve(bytes ? bytes : 1));
    *dev = buf.ptr;
    return UCFP_OK;
}

int copy_back(ucfp_lane *ln, void *user, const void *dev, size_t bytes) {
    if (bytes) UCFP_CUDA_TRY(cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, ln->stream));
    return UCFP_OK;
}

// End of a call: pooled mode always completes the work (the lane goes back to the pool, its scratch may be reused by
// another thread at once); shared-stream mode synchronises only when the caller reads host memory afterwards.
int finish_call(ucfp_lane *ln, bool host_outputs) {
    if (host_outputs || !ln->owner->shared_stream) UCFP_CUDA_TRY(cudaStreamSynchronize(ln->stream));
    return UCFP_OK;
}

size_t row_bytes(const ucfp_corpus *c) {
    switch (c->kind) {
        case UCFP_KIND_HAMMING64: return 8;
        case UCFP_KIND_MINHASH128: return 1024;
        case UCFP_KIND_COSINE: return 4 * (size_t)c->dim;
        case UCFP_KIND_MULTIHASH: return sizeof(ucfp_image_hashes);
    }
    return 0;
}

int after_append(ucfp_lane *ln, ucfp_corpus *c, uint64_t first, uint64_t n) {
    if (c->kind == UCFP_KIND_HAMMING64) return hamming_on_append(ln, c, first, n);
    if (c->kind == UCFP_KIND_MINHASH128) return jaccard_on_append(ln, c, first, n);
    if (c->kind == UCFP_KIND_COSINE) return cosine_on_append(ln, c, first, n);
    if (c->kind == UCFP_KIND_MULTIHASH) return multihash_on_append(ln, c, first, n);
    return UCFP_OK;
}

}  // namespace ucfp

using namespace ucfp;

extern "C" {

int ucfp_abi_version(void) { return UCFP_ABI_VERSION; }

const char *ucfp_last_error(void) { return g_err; }

int ucfp_init(int device, ucfp_ctx **out) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(out != nullptr, UCFP_E_INVALID, "ucfp_init: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libucfp_cuda has no CPU fallback", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        return UCFP_E_CUDA;
    }
    UCFP_REQUIRE(device >= 0 && device < ndev, UCFP_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    cudaDeviceProp prop;
    UCFP_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    UCFP_REQUIRE(prop.major == 10, UCFP_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                 prop.major, prop.minor);
    ucfp_ctx *ctx = new (std::nothrow) ucfp_ctx();
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_OOM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    DeviceGuard dg(device);
    ctx->lanes[0] = lane_create(ctx);
    if (!ctx->lanes[0]) { delete ctx; set_error("cudaStreamCreate failed"); return UCFP_E_CUDA; }
    ctx->n_lanes = 1;
    // per-device kernel attributes and occupancies, once
    int rc = hamming_device_init(ctx);
    if (rc == UCFP_OK) rc = jaccard_device_init(ctx);
    if (rc == UCFP_OK) rc = cosine_device_init(ctx);
    if (rc == UCFP_OK) rc = image_device_init(ctx);
    if (rc == UCFP_OK) rc = merge_device_init(ctx);
    if (rc == UCFP_OK) rc = corpus_device_init(ctx);
    if (rc != UCFP_OK) { lane_destroy(ctx->lanes[0]); delete ctx; return rc; }
    *out = ctx;
    return UCFP_OK;
    UCFP_API_END
}

void ucfp_destroy(ucfp_ctx *ctx) {
    if (!ctx) return;
    try {
        DeviceGuard dg(ctx->device);
        for (int i = 0; i < ctx->n_lanes; ++i) lane_destroy(ctx->lanes[i]);
        image_cache_destroy(ctx);
        jpeg_destroy(ctx);
        for (auto &r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        delete ctx;
    } catch (...) {
    }
}

int ucfp_ctx_set_stream(ucfp_ctx *ctx, void *cuda_stream) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    UCFP_LEASE(ctx);   // waits for running calls of shared-stream mode; pooled calls in flight finish on their own lanes
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->shared_stream = true;
    ctx->user_stream = static_cast<cudaStream_t>(cuda_stream);
    return UCFP_OK;
    UCFP_API_END
}

int ucfp_ctx_reset_stream(ucfp_ctx *ctx) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    UCFP_LEASE(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->shared_stream = false;
    ctx->user_stream = nullptr;
    return UCFP_OK;
    UCFP_API_END
}

int ucfp_ctx_synchronize(ucfp_ctx *ctx) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    DeviceGuard dg(ctx->device);
    UCFP_REQUIRE(dg.ok, UCFP_E_CUDA, "cudaSetDevice(%d) failed", ctx->device);
    cudaStream_t streams[kUcfpMaxLanes + 1];
    int n = 0;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        for (int i = 0; i < ctx->n_lanes; ++i) streams[n++] = ctx->lanes[i]->own_stream;
        if (ctx->shared_stream) streams[n++] = ctx->user_stream;
    }
    for (int i = 0; i < n; ++i) UCFP_CUDA_TRY(cudaStreamSynchronize(streams[i]));
    return UCFP_OK;
    UCFP_API_END
}

uint64_t ucfp_ctx_kernel_launches(const ucfp_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

static void prof_clear(ucfp_ctx *ctx) {
    for (auto &r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    ctx->prof.clear();
}

int ucfp_ctx_profile_begin(ucfp_ctx *ctx) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    prof_clear(ctx);
    ctx->profiling.store(true);
    return UCFP_OK;
    UCFP_API_END
}

static int prof_sum(ucfp_ctx *ctx, int kernel_class, double *kernel_ms, double *alg_units, uint64_t *launches, bool end) {
    UCFP_TRY(ucfp_ctx_synchronize(ctx));
    DeviceGuard dg(ctx->device);
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    double ms = 0, units = 0; uint64_t n = 0;
    for (auto &r : ctx->prof) {
        if (r.kind != kernel_class) continue;
        float t = 0;
        UCFP_CUDA_TRY(cudaEventElapsedTime(&t, r.a, r.b));
        ms += t; units += r.units; n++;
    }
    if (kernel_ms) *kernel_ms = ms;
    if (alg_units) *alg_units = units;
    if (launches) *launches = n;
    if (end) { ctx->profiling.store(false); prof_clear(ctx); }
    return UCFP_OK;
}

int ucfp_ctx_profile_read(ucfp_ctx *ctx, int kernel_class, double *kernel_ms, double *alg_units, uint64_t *launches) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    return prof_sum(ctx, kernel_class, kernel_ms, alg_units, launches, false);
    UCFP_API_END
}

int ucfp_ctx_profile_end(ucfp_ctx *ctx, int kernel_class, double *kernel_ms, double *alg_units, uint64_t *launches) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    return prof_sum(ctx, kernel_class, kernel_ms, alg_units, launches, true);
    UCFP_API_END
}

// ---- corpus -------------------------------------------------------------------------------

static void corpus_free_arrays(ucfp_corpus *c) {
    if (c->coarse) { c->coarse->ids = nullptr; /* borrowed from c */ corpus_free_arrays(c->coarse); delete c->coarse; c->coarse = nullptr; }
    if (c->rows) cudaFree(c->rows);
    if (c->ids) cudaFree(c->ids);
    if (c->ham_ops) cudaFree(c->ham_ops);
    if (c->mh_sketch) cudaFree(c->mh_sketch);
    if (c->cos_bf16) cudaFree(c->cos_bf16);
    if (c->cos_inv_norm) cudaFree(c->cos_inv_norm);
    c->rows = nullptr; c->ids = nullptr; c->ham_ops = nullptr; c->mh_sketch = nullptr; c->cos_bf16 = nullptr; c->cos_inv_norm = nullptr;
}

int ucfp_corpus_create(ucfp_ctx *ctx, int kind, uint32_t dim, uint64_t capacity, ucfp_corpus **out) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(out != nullptr, UCFP_E_INVALID, "ucfp_corpus_create: out is NULL");
    *out = nullptr;
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    UCFP_LEASE(ctx);
    UCFP_REQUIRE(kind == UCFP_KIND_HAMMING64 || kind == UCFP_KIND_MINHASH128 || kind == UCFP_KIND_COSINE || kind == UCFP_KIND_MULTIHASH,
                 UCFP_E_INVALID, "unknown corpus kind %d", kind);
    UCFP_REQUIRE(capacity > 0, UCFP_E_INVALID, "capacity must be > 0");
    if (kind == UCFP_KIND_COSINE) UCFP_REQUIRE(dim > 0 && dim <= 4096, UCFP_E_INVALID, "cosine dim must be in 1..4096 (got %u)", dim);
    ucfp_corpus *c = new (std::nothrow) ucfp_corpus();
    UCFP_REQUIRE(c != nullptr, UCFP_E_OOM, "out of host memory");
    c->ctx = ctx; c->kind = kind; c->dim = kind == UCFP_KIND_COSINE ? dim : 0;
    int rc = corpus_alloc_arrays(lane, c, capacity);
    if (rc != UCFP_OK) { delete c; return rc; }
    UCFP_TRY(finish_call(lane, false));
    *out = c;
    return UCFP_OK;
    UCFP_API_END
}

void ucfp_corpus_destroy(ucfp_corpus *c) {
    if (!c) return;
    try {
        {
            std::unique_lock<std::shared_mutex> wl(c->rw);   // waits for scans in flight
            ucfp_ctx_synchronize(c->ctx);
            DeviceGuard dg(c->ctx->device);
            corpus_free_arrays(c);
        }
        delete c;
    } catch (...) {
    }
}

uint64_t ucfp_corpus_size(const ucfp_corpus *c) { return c ? c->size : 0; }
uint64_t ucfp_corpus_capacity(const ucfp_corpus *c) { return c ? c->capacity : 0; }

void *ucfp_corpus_device_rows(ucfp_corpus *c) { return c ? c->rows : nullptr; }

int ucfp_corpus_set_id_base(ucfp_corpus *c, uint64_t id_base) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    std::unique_lock<std::shared_mutex> wl(c->rw);
    c->id_base = id_base;
    return UCFP_OK;
    UCFP_API_END
}

int ucfp_corpus_clear(ucfp_corpus *c) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    std::unique_lock<std::shared_mutex> wl(c->rw);
    c->size = 0;
    c->id_mode = 0;
    return UCFP_OK;
    UCFP_API_END
}

int ucfp_corpus_refresh(ucfp_corpus *c) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    std::unique_lock<std::shared_mutex> wl(c->rw);
    UCFP_LEASE(c->ctx);
    UCFP_TRY(after_append(lane, c, 0, c->size));
    return finish_call(lane, false);
    UCFP_API_END
}

static int append_common(ucfp_corpus *c, const uint64_t *ids, const void *src, uint64_t src_stride, uint64_t field_offset, uint64_t n,
                         bool strided) {
    std::unique_lock<std::shared_mutex> wl(c->rw);
    UCFP_LEASE(c->ctx);
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(src != nullptr, UCFP_E_INVALID, "rows is NULL");
    const size_t rb = row_bytes(c);
    const bool bundles = strided && c->kind == UCFP_KIND_MULTIHASH;   // records are 536-byte MultiHashFingerprints starting at field_offset
    if (strided)
        UCFP_REQUIRE(src_stride >= field_offset + (bundles ? 536 : rb), UCFP_E_INVALID, "record stride %llu cannot hold a %zu-byte field at offset %llu",
                     (unsigned long long)src_stride, bundles ? (size_t)536 : rb, (unsigned long long)field_offset);
    UCFP_REQUIRE(c->size + n <= c->capacity, UCFP_E_CAPACITY, "append of %llu rows exceeds capacity %llu (size %llu)",
                 (unsigned long long)n, (unsigned long long)c->capacity, (unsigned long long)c->size);
    const int mode = ids ? 1 : 2;
    UCFP_REQUIRE(c->id_mode == 0 || c->id_mode == mode, UCFP_E_STATE, "corpus mixes explicit and implicit record ids");
    cudaStream_t st = lane->stream;
    if (mode == 1 && !c->ids) UCFP_CUDA_TRY(cudaMalloc((void **)&c->ids, 8 * (c->capacity + 16)));
    const bool dev_src = classify(src) == Mem::Device;
    const cudaMemcpyKind kr = dev_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    char *dst = static_cast<char *>(c->rows) + rb * c->size;
    if (bundles) {   // exact[32] | ahash:ImageFingerprint | phash | dhash, each ImageFingerprint = exact[32] | global | 16 blocks: three 136-byte runs
        for (int a = 0; a < 3; ++a)
            UCFP_CUDA_TRY(cudaMemcpy2DAsync(dst + 136 * a, rb, static_cast<const char *>(src) + field_offset + 64 + 168 * a, src_stride, 136, n, kr, st));
    } else if (strided)   // a pitched copy gathers the field of every record: source pitch = record stride, width = one row
        UCFP_CUDA_TRY(cudaMemcpy2DAsync(dst, rb, static_cast<const char *>(src) + field_offset, src_stride, rb, n, kr, st));
    else
        UCFP_CUDA_TRY(cudaMemcpyAsync(dst, src, rb * n, kr, st));
    bool host_ids = false;
    if (mode == 1) {
        host_ids = classify(ids) != Mem::Device;
        UCFP_CUDA_TRY(cudaMemcpyAsync(c->ids + c->size, ids, 8 * n, host_ids ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    }
    UCFP_TRY(after_append(lane, c, c->size, n));
    // host sources may be reused by the caller as soon as we return
    UCFP_TRY(finish_call(lane, !dev_src || host_ids));
    c->id_mode = mode;
    c->size += n;
    return UCFP_OK;
}

int ucfp_corpus_append(ucfp_corpus *c, const uint64_t *ids, const void *rows, uint64_t n) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    return append_common(c, ids, rows, 0, 0, n, false);
    UCFP_API_END
}

int ucfp_corpus_append_strided(ucfp_corpus *c, const uint64_t *ids, const void *records, uint64_t record_stride,
                               uint64_t field_offset, uint64_t n) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    return append_common(c, ids, records, record_stride, field_offset, n, true);
    UCFP_API_END
}

int ucfp_corpus_append_synthetic(ucfp_corpus *c, uint64_t seed, uint64_t start_row, uint64_t n) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    std::unique_lock<std::shared_mutex> wl(c->rw);
    UCFP_LEASE(c->ctx);
    UCFP_REQUIRE(c->kind == UCFP_KIND_HAMMING64 || c->kind == UCFP_KIND_MINHASH128, UCFP_E_UNSUPPORTED,
                 "synthetic rows exist for HAMMING64 and MINHASH128 corpora only");
    UCFP_REQUIRE(c->size + n <= c->capacity, UCFP_E_CAPACITY, "append of %llu rows exceeds capacity", (unsigned long long)n);
    UCFP_REQUIRE(c->id_mode == 0 || c->id_mode == 2, UCFP_E_STATE, "corpus mixes explicit and implicit record ids");
    uint64_t wpr = c->kind == UCFP_KIND_HAMMING64 ? 1 : 128;
    UCFP_TRY(synth_fill_u64(lane, static_cast<uint64_t *>(c->rows) + c->size * wpr, n * wpr, seed, start_row * wpr));
    UCFP_TRY(after_append(lane, c, c->size, n));
    UCFP_TRY(finish_call(lane, false));
    c->id_mode = 2;
    c->size += n;
    return UCFP_OK;
    UCFP_API_END
}

int ucfp_corpus_reserve(ucfp_corpus *c, uint64_t capacity) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    std::unique_lock<std::shared_mutex> wl(c->rw);
    UCFP_LEASE(c->ctx);
    if (capacity <= c->capacity) return UCFP_OK;
    UCFP_TRY(corpus_grow(lane, c, capacity));
    return finish_call(lane, true);
    UCFP_API_END
}

int ucfp_corpus_delete(ucfp_corpus *c, const uint64_t *ids, uint64_t n, uint64_t *n_removed) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    if (n_removed) *n_removed = 0;
    std::unique_lock<std::shared_mutex> wl(c->rw);
    UCFP_LEASE(c->ctx);
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(ids != nullptr, UCFP_E_INVALID, "ids is NULL");
    UCFP_TRY(corpus_delete_ids(lane, c, ids, n, n_removed));
    return finish_call(lane, true);
    UCFP_API_END
}

int ucfp_corpus_upsert(ucfp_corpus *c, const uint64_t *ids, const void *rows, uint64_t n, uint64_t *n_replaced) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    if (n_replaced) *n_replaced = 0;
    std::unique_lock<std::shared_mutex> wl(c->rw);
    UCFP_LEASE(c->ctx);
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(ids != nullptr && rows != nullptr, UCFP_E_INVALID, "ids or rows is NULL");
    UCFP_TRY(corpus_upsert_rows(lane, c, ids, rows, n, n_replaced));
    return finish_call(lane, true);
    UCFP_API_END
}

}  // extern "C"

// ---- scans --------------------------------------------------------------------------------
namespace ucfp {

// Shared by the scan entry points and the batcher (batcher.cu): one batch of queries against one corpus.
int run_scan_any(ucfp_corpus *c, int want_kind, const void *queries, size_t nq, size_t k, uint64_t *ids_out, void *keys_out) {
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    std::shared_lock<std::shared_mutex> rl(c->rw);
    UCFP_LEASE(c->ctx);
    UCFP_REQUIRE(c->kind == want_kind, UCFP_E_STATE, "corpus kind %d cannot serve this scan (needs kind %d)", c->kind, want_kind);
    if (nq == 0 || k == 0) return UCFP_OK;
    UCFP_REQUIRE(queries && ids_out && keys_out, UCFP_E_INVALID, "NULL query or output buffer");
    const size_t q_bytes = nq * row_bytes(c);
    const void *q_dev = nullptr;
    UCFP_TRY(stage_in(lane, lane->q_dev, queries, q_bytes, &q_dev));
    void *ids_dev = nullptr, *keys_dev = nullptr;
    bool ids_host = false, keys_host = false;
    UCFP_TRY(stage_out(lane->out_ids_dev, ids_out, 8 * nq * k, &ids_dev, &ids_host));
    UCFP_TRY(stage_out(lane->out_keys_dev, keys_out, 4 * nq * k, &keys_dev, &keys_host));
    UCFP_TRY(stats_reset(lane));
    if (want_kind == UCFP_KIND_HAMMING64)
        UCFP_TRY(hamming_scan(lane, c, static_cast<const uint64_t *>(q_dev), nq, k, static_cast<uint64_t *>(ids_dev), static_cast<uint32_t *>(keys_dev)));
    else if (want_kind == UCFP_KIND_MINHASH128)
        UCFP_TRY(jaccard_scan(lane, c, static_cast<const uint64_t *>(q_dev), nq, k, static_cast<uint64_t *>(ids_dev), static_cast<uint32_t *>(keys_dev)));
    else
        UCFP_TRY(cosine_scan(lane, c, static_cast<const float *>(q_dev), nq, k, static_cast<uint64_t *>(ids_dev), static_cast<float *>(keys_dev)));
    if (ids_host) UCFP_TRY(copy_back(lane, ids_out, ids_dev, 8 * nq * k));
    if (keys_host) UCFP_TRY(copy_back(lane, keys_out, keys_dev, 4 * nq * k));
    {
        std::lock_guard<std::mutex> lk(c->ctx->mu);
        for (int i = 0; i < c->ctx->n_lanes; ++i) if (c->ctx->lanes[i] == lane) c->ctx->last_scan_lane = i;
    }
    return finish_call(lane, ids_host || keys_host);
}

}  // namespace ucfp

extern "C" {

int ucfp_scan_hamming(ucfp_corpus *c, const uint64_t *queries, size_t nq, size_t k, uint64_t *ids_out, uint32_t *dist_out) {
    UCFP_API_BEGIN
    return run_scan_any(c, UCFP_KIND_HAMMING64, queries, nq, k, ids_out, dist_out);
    UCFP_API_END
}

int ucfp_scan_jaccard(ucfp_corpus *c, const uint64_t *queries, size_t nq, size_t k, uint64_t *ids_out, uint32_t *matches_out) {
    UCFP_API_BEGIN
    return run_scan_any(c, UCFP_KIND_MINHASH128, queries, nq, k, ids_out, matches_out);
    UCFP_API_END
}

int ucfp_scan_cosine(ucfp_corpus *c, const float *queries, size_t nq, size_t k, uint64_t *ids_out, float *score_out) {
    UCFP_API_BEGIN
    return run_scan_any(c, UCFP_KIND_COSINE, queries, nq, k, ids_out, score_out);
    UCFP_API_END
}

int ucfp_ctx_last_scan_stats(ucfp_ctx *ctx, uint64_t *queries_recomputed, uint64_t *max_list_fill) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    if (queries_recomputed) *queries_recomputed = 0;
    if (max_list_fill) *max_list_fill = 0;
    UCFP_TRY(ucfp_ctx_synchronize(ctx));
    DeviceGuard dg(ctx->device);
    void *stats = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        stats = ctx->lanes[ctx->last_scan_lane]->stats.ptr;
    }
    if (!stats) return UCFP_OK;
    uint64_t h[2] = {0, 0};
    UCFP_CUDA_TRY(cudaMemcpy(h, stats, 16, cudaMemcpyDeviceToHost));
    if (queries_recomputed) *queries_recomputed = h[0];
    if (max_list_fill) *max_list_fill = h[1];
    return UCFP_OK;
    UCFP_API_END
}

int ucfp_ctx_last_scan_exact_selects(ucfp_ctx *ctx, uint64_t *queries) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    UCFP_REQUIRE(queries != nullptr, UCFP_E_INVALID, "NULL output");
    *queries = 0;
    UCFP_TRY(ucfp_ctx_synchronize(ctx));
    DeviceGuard dg(ctx->device);
    void *stats = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        stats = ctx->lanes[ctx->last_scan_lane]->stats.ptr;
    }
    if (!stats) return UCFP_OK;
    UCFP_CUDA_TRY(cudaMemcpy(queries, static_cast<const uint64_t *>(stats) + 2, 8, cudaMemcpyDeviceToHost));
#ifdef UCFP_DEBUG_RESCAN   // developer build: longest re-scan list and number of re-scan compactions in the upper bits
    uint64_t dbg[2] = {0, 0};
    UCFP_CUDA_TRY(cudaMemcpy(dbg, static_cast<const uint64_t *>(stats) + 3, 16, cudaMemcpyDeviceToHost));
    *queries |= dbg[0] << 16 | dbg[1] << 48;
#endif
    return UCFP_OK;
    UCFP_API_END
}

int ucfp_ctx_last_scan_fallbacks(ucfp_ctx *ctx, uint64_t *queries_recomputed) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    UCFP_REQUIRE(queries_recomputed != nullptr, UCFP_E_INVALID, "NULL output");
    *queries_recomputed = 0;
    UCFP_TRY(ucfp_ctx_synchronize(ctx));
    DeviceGuard dg(ctx->device);
    void *stats = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        stats = ctx->lanes[ctx->last_scan_lane]->stats.ptr;
    }
    if (!stats) return UCFP_OK;
    UCFP_CUDA_TRY(cudaMemcpy(queries_recomputed, stats, 8, cudaMemcpyDeviceToHost));
    return UCFP_OK;
    UCFP_API_END
}

}  // extern "C"

template <typename Key, typename MergeFn>
static int run_merge(ucfp_ctx *ctx, const uint64_t *ids_in, const Key *keys_in, size_t parts, size_t nq, size_t k,
                     uint64_t *ids_out, Key *keys_out, MergeFn merge) {
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    UCFP_LEASE(ctx);
    if (nq == 0 || k == 0 || parts == 0) return UCFP_OK;
    UCFP_REQUIRE(ids_in && keys_in && ids_out && keys_out, UCFP_E_INVALID, "NULL buffer");
    size_t n_in = parts * nq * k;
    const void *ids_in_dev = nullptr, *keys_in_dev = nullptr;
    UCFP_TRY(stage_in(lane, lane->cand, ids_in, 8 * n_in, &ids_in_dev));
    UCFP_TRY(stage_in(lane, lane->misc, keys_in, sizeof(Key) * n_in, &keys_in_dev));
    void *ids_dev = nullptr, *keys_dev = nullptr;
    bool ids_host = false, keys_host = false;
    UCFP_TRY(stage_out(lane->out_ids_dev, ids_out, 8 * nq * k, &ids_dev, &ids_host));
    UCFP_TRY(stage_out(lane->out_keys_dev, keys_out, sizeof(Key) * nq * k, &keys_dev, &keys_host));
    UCFP_TRY(merge(lane, static_cast<const uint64_t *>(ids_in_dev), static_cast<const Key *>(keys_in_dev),
                   static_cast<uint64_t *>(ids_dev), static_cast<Key *>(keys_dev)));
    if (ids_host) UCFP_TRY(copy_back(lane, ids_out, ids_dev, 8 * nq * k));
    if (keys_host) UCFP_TRY(copy_back(lane, keys_out, keys_dev, sizeof(Key) * nq * k));
    return finish_call(lane, ids_host || keys_host);
}

// shared body of the two image entry points (pixels_mem: see image_hash_batch)
static int image_batch_call(ucfp_ctx *ctx, const ucfp_image_desc *imgs, size_t n, uint32_t algo_mask, ucfp_image_hashes *out, int32_t *status,
                            int pixels_mem) {
    using namespace ucfp;
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    UCFP_LEASE(ctx);
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(imgs && out, UCFP_E_INVALID, "NULL image descriptors or output");
    UCFP_REQUIRE((algo_mask & ~UCFP_ALGO_MULTI) == 0 && algo_mask != 0, UCFP_E_INVALID, "bad algo_mask 0x%x", algo_mask);
    void *out_dev = nullptr;
    bool out_host = false;
    UCFP_TRY(stage_out(lane->img_out_dev, out, sizeof(ucfp_image_hashes) * n, &out_dev, &out_host));
    // per-image status: straight into the caller's array when it is host memory, else through the lane's pinned staging
    const bool status_dev = status && classify(status) == Mem::Device;
    int32_t *st_host = status;
    if (!status || status_dev) {
        UCFP_TRY(lane->pin_b.reserve(4 * n));
        st_host = lane->pin_b.as<int32_t>();
    }
    UCFP_TRY(image_hash_batch(lane, imgs, n, algo_mask, static_cast<ucfp_image_hashes *>(out_dev), st_host, pixels_mem));
    if (out_host) UCFP_TRY(copy_back(lane, out, out_dev, sizeof(ucfp_image_hashes) * n));
    if (status_dev) UCFP_CUDA_TRY(cudaMemcpyAsync(status, st_host, 4 * n, cudaMemcpyHostToDevice, lane->stream));
    return finish_call(lane, out_host || status_dev);
}

extern "C" {

int ucfp_merge_topk_u32(ucfp_ctx *ctx, const uint64_t *ids_in, const uint32_t *keys_in, size_t parts, size_t nq, size_t k,
                        int descending, uint64_t *ids_out, uint32_t *keys_out) {
    UCFP_API_BEGIN
    return run_merge<uint32_t>(ctx, ids_in, keys_in, parts, nq, k, ids_out, keys_out,
                               [&](ucfp_lane *ln, const uint64_t *ii, const uint32_t *ki, uint64_t *io, uint32_t *ko) {
                                   return merge_u32(ln, ii, ki, parts, nq, k, descending, io, ko);
                               });
    UCFP_API_END
}

int ucfp_merge_topk_f32(ucfp_ctx *ctx, const uint64_t *ids_in, const float *scores_in, size_t parts, size_t nq, size_t k,
                        uint64_t *ids_out, float *scores_out) {
    UCFP_API_BEGIN
    return run_merge<float>(ctx, ids_in, scores_in, parts, nq, k, ids_out, scores_out,
                            [&](ucfp_lane *ln, const uint64_t *ii, const float *ki, uint64_t *io, float *ko) {
                                return merge_f32(ln, ii, ki, parts, nq, k, io, ko);
                            });
    UCFP_API_END
}

// ---- image hashing ------------------------------------------------------------------------

int ucfp_image_hash_batch(ucfp_ctx *ctx, const ucfp_image_desc *imgs, size_t n, uint32_t algo_mask, ucfp_image_hashes *out,
                          int32_t *status) {
    UCFP_API_BEGIN
    return image_batch_call(ctx, imgs, n, algo_mask, out, status, -1);
    UCFP_API_END
}

int ucfp_image_hash_uniform(ucfp_ctx *ctx, const uint8_t *pixels, size_t n, uint32_t width, uint32_t height, uint64_t row_stride,
                            uint64_t image_stride, uint32_t algo_mask, ucfp_image_hashes *out) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(pixels && out, UCFP_E_INVALID, "NULL pixels or output");
    UCFP_REQUIRE(width >= 4 && height >= 4 && row_stride >= 3ull * width && image_stride >= row_stride * (height - 1) + 3ull * width,
                 UCFP_E_INVALID, "bad image geometry %ux%u stride %llu/%llu", width, height, (unsigned long long)row_stride,
                 (unsigned long long)image_stride);
    std::vector<ucfp_image_desc> d(n);
    for (size_t i = 0; i < n; ++i) d[i] = ucfp_image_desc{pixels + i * image_stride, width, height, row_stride};
    std::vector<int32_t> st(n, 0);
    // one buffer: ask once where it lives (its first and last byte), not once per image
    const uint8_t *last = pixels + (n - 1) * image_stride + row_stride * (height - 1) + 3ull * width - 1;
    const Mem m0 = classify(pixels), m1 = classify(last);
    UCFP_REQUIRE(m0 == m1, UCFP_E_INVALID, "the pixel buffer starts and ends in different kinds of memory");
    int rc = image_batch_call(ctx, d.data(), n, algo_mask, out, st.data(), m0 == Mem::Device ? 1 : 0);
    if (rc != UCFP_OK) return rc;
    for (size_t i = 0; i < n; ++i)
        if (st[i] != UCFP_OK) { set_error("image %zu failed with status %d", i, st[i]); return st[i]; }
    return UCFP_OK;
    UCFP_API_END
}

}  // extern "C"
