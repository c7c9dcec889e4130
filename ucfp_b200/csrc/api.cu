// api.cu -- the extern "C" surface of libucfp_cuda.so (declared in include/ucfp_cuda.h):
// context and corpus lifetime, host/device staging, dispatch into the per-path kernels.
#include <stdarg.h>

#include "common.cuh"

namespace ucfp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) { ok = false; cudaGetLastError(); }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Makes `user` (host or device) readable on the device.  Host data is copied into `buf`.
int stage_in(ucfp_ctx *ctx, DevBuf &buf, const void *user, size_t bytes, const void **dev) {
    if (bytes == 0) { *dev = buf.ptr; return UCFP_OK; }
    if (classify(user) == Mem::Device) { *dev = user; return UCFP_OK; }
    UCFP_TRY(buf.reserve(bytes));
    UCFP_CUDA_TRY(cudaMemcpyAsync(buf.ptr, user, bytes, cudaMemcpyHostToDevice, ctx->stream));
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

int copy_back(ucfp_ctx *ctx, void *user, const void *dev, size_t bytes) {
    if (bytes) UCFP_CUDA_TRY(cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return UCFP_OK;
}

size_t row_bytes(const ucfp_corpus *c) {
    switch (c->kind) {
        case UCFP_KIND_HAMMING64: return 8;
        case UCFP_KIND_MINHASH128: return 1024;
        case UCFP_KIND_COSINE: return 4 * (size_t)c->dim;
    }
    return 0;
}

}  // namespace
}  // namespace ucfp

using namespace ucfp;

#define UCFP_GUARD(ctxp)                                                              \
    UCFP_REQUIRE((ctxp) != nullptr, UCFP_E_INVALID, "null context");                  \
    DeviceGuard _dg((ctxp)->device);                                                  \
    UCFP_REQUIRE(_dg.ok, UCFP_E_CUDA, "cudaSetDevice(%d) failed", (ctxp)->device);    \
    std::lock_guard<std::mutex> _lk((ctxp)->mu)

extern "C" {

int ucfp_abi_version(void) { return UCFP_ABI_VERSION; }

const char *ucfp_last_error(void) { return g_err; }

int ucfp_init(int device, ucfp_ctx **out) {
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
    cudaError_t se = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) { delete ctx; set_error("cudaStreamCreate failed: %s", cudaGetErrorString(se)); return UCFP_E_CUDA; }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return UCFP_OK;
}

void ucfp_destroy(ucfp_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard dg(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->q_dev, &ctx->out_ids_dev, &ctx->out_keys_dev, &ctx->cand, &ctx->cand_count, &ctx->qstate, &ctx->flags,
                      &ctx->misc, &ctx->img_desc_dev, &ctx->img_out_dev, &ctx->img_status_dev, &ctx->img_tables_dev, &ctx->img_stage_dev, &ctx->stats};
    for (DevBuf *b : bufs) b->release();
    ctx->pin_a.release(); ctx->pin_b.release();
    image_cache_destroy(ctx);
    for (auto &r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int ucfp_ctx_set_stream(ucfp_ctx *ctx, void *cuda_stream) {
    UCFP_GUARD(ctx);
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return UCFP_OK;
}

int ucfp_ctx_reset_stream(ucfp_ctx *ctx) {
    UCFP_GUARD(ctx);
    ctx->stream = ctx->own_stream;
    return UCFP_OK;
}

int ucfp_ctx_synchronize(ucfp_ctx *ctx) {
    UCFP_GUARD(ctx);
    UCFP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return UCFP_OK;
}

uint64_t ucfp_ctx_kernel_launches(const ucfp_ctx *ctx) { return ctx ? ctx->launches : 0; }

static void prof_clear(ucfp_ctx *ctx) {
    for (auto &r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    ctx->prof.clear();
}

int ucfp_ctx_profile_begin(ucfp_ctx *ctx) {
    UCFP_GUARD(ctx);
    prof_clear(ctx);
    ctx->profiling = true;
    return UCFP_OK;
}

int ucfp_ctx_profile_read(ucfp_ctx *ctx, int kernel_class, double *kernel_ms, double *alg_units, uint64_t *launches) {
    UCFP_GUARD(ctx);
    UCFP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
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
    return UCFP_OK;
}

int ucfp_ctx_profile_end(ucfp_ctx *ctx, int kernel_class, double *kernel_ms, double *alg_units, uint64_t *launches) {
    UCFP_GUARD(ctx);
    UCFP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
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
    ctx->profiling = false;
    prof_clear(ctx);
    return UCFP_OK;
}

// ---- corpus -------------------------------------------------------------------------------

int ucfp_corpus_create(ucfp_ctx *ctx, int kind, uint32_t dim, uint64_t capacity, ucfp_corpus **out) {
    UCFP_REQUIRE(out != nullptr, UCFP_E_INVALID, "ucfp_corpus_create: out is NULL");
    *out = nullptr;
    UCFP_GUARD(ctx);
    UCFP_REQUIRE(kind == UCFP_KIND_HAMMING64 || kind == UCFP_KIND_MINHASH128 || kind == UCFP_KIND_COSINE, UCFP_E_INVALID,
                 "unknown corpus kind %d", kind);
    UCFP_REQUIRE(capacity > 0, UCFP_E_INVALID, "capacity must be > 0");
    if (kind == UCFP_KIND_COSINE) UCFP_REQUIRE(dim > 0 && dim <= 4096, UCFP_E_INVALID, "cosine dim must be in 1..4096 (got %u)", dim);
    ucfp_corpus *c = new (std::nothrow) ucfp_corpus();
    UCFP_REQUIRE(c != nullptr, UCFP_E_OOM, "out of host memory");
    c->ctx = ctx; c->kind = kind; c->dim = kind == UCFP_KIND_COSINE ? dim : 0; c->capacity = capacity;
    size_t rb = row_bytes(c);
    // +16 rows of slack so that vector loads of the last partial tile never leave the allocation
    cudaError_t e = cudaMalloc(&c->rows, rb * (capacity + 16));
    if (e == cudaSuccess && kind == UCFP_KIND_MINHASH128) e = cudaMalloc((void **)&c->mh_sketch, 2 * 128 * ((capacity + 31) / 32 * 32 + 512));  // two planes (byte 0, byte 1 of every slot); whole 256-row tiles stay readable
    if (e == cudaSuccess && kind == UCFP_KIND_COSINE) {
        c->dim_pad = (dim + 63) / 64 * 64;
        e = cudaMalloc(&c->cos_bf16, 2 * (size_t)c->dim_pad * (capacity + 256));
        if (e == cudaSuccess) e = cudaMalloc((void **)&c->cos_inv_norm, 4 * (capacity + 256));
    }
    if (e == cudaSuccess && kind == UCFP_KIND_HAMMING64) {
        // Operand rows of the tensor-core scan, 32 B per code on top of the 8 B code.  Optional: without them (allocation
        // refused) the scan expands the codes on the fly in its producer warps, slower for 64-512-query batches.
        const size_t ops_bytes = 64 * ((capacity + 1) / 2 + 512);   // whole 256-row stages stay readable
        if (cudaMalloc((void **)&c->ham_ops, ops_bytes) != cudaSuccess) { cudaGetLastError(); c->ham_ops = nullptr; }
        else if (cudaMemsetAsync(c->ham_ops, 0, ops_bytes, ctx->stream) != cudaSuccess) { cudaGetLastError(); cudaFree(c->ham_ops); c->ham_ops = nullptr; }
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("corpus allocation of %llu rows failed: %s", (unsigned long long)capacity, cudaGetErrorString(e));
        if (c->rows) cudaFree(c->rows);
        if (c->mh_sketch) cudaFree(c->mh_sketch);
        if (c->cos_bf16) cudaFree(c->cos_bf16);
        if (c->cos_inv_norm) cudaFree(c->cos_inv_norm);
        delete c;
        return UCFP_E_OOM;
    }
    *out = c;
    return UCFP_OK;
}

void ucfp_corpus_destroy(ucfp_corpus *c) {
    if (!c) return;
    {
        DeviceGuard dg(c->ctx->device);
        std::lock_guard<std::mutex> lk(c->ctx->mu);
        cudaStreamSynchronize(c->ctx->stream);
        if (c->rows) cudaFree(c->rows);
        if (c->ids) cudaFree(c->ids);
        if (c->ham_ops) cudaFree(c->ham_ops);
        if (c->mh_sketch) cudaFree(c->mh_sketch);
        if (c->cos_bf16) cudaFree(c->cos_bf16);
        if (c->cos_inv_norm) cudaFree(c->cos_inv_norm);
    }
    delete c;
}

uint64_t ucfp_corpus_size(const ucfp_corpus *c) { return c ? c->size : 0; }

void *ucfp_corpus_device_rows(ucfp_corpus *c) { return c ? c->rows : nullptr; }

int ucfp_corpus_set_id_base(ucfp_corpus *c, uint64_t id_base) {
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    UCFP_GUARD(c->ctx);
    c->id_base = id_base;
    return UCFP_OK;
}

int ucfp_corpus_clear(ucfp_corpus *c) {
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    UCFP_GUARD(c->ctx);
    c->size = 0;
    c->id_mode = 0;
    return UCFP_OK;
}

static int after_append(ucfp_corpus *c, uint64_t first, uint64_t n) {
    if (c->kind == UCFP_KIND_HAMMING64) return hamming_on_append(c, first, n);
    if (c->kind == UCFP_KIND_MINHASH128) return jaccard_on_append(c, first, n);
    if (c->kind == UCFP_KIND_COSINE) return cosine_on_append(c, first, n);
    return UCFP_OK;
}

int ucfp_corpus_refresh(ucfp_corpus *c) {
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    UCFP_GUARD(c->ctx);
    return after_append(c, 0, c->size);
}

int ucfp_corpus_append(ucfp_corpus *c, const uint64_t *ids, const void *rows, uint64_t n) {
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    UCFP_GUARD(c->ctx);
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(rows != nullptr, UCFP_E_INVALID, "rows is NULL");
    UCFP_REQUIRE(c->size + n <= c->capacity, UCFP_E_CAPACITY, "append of %llu rows exceeds capacity %llu (size %llu)",
                 (unsigned long long)n, (unsigned long long)c->capacity, (unsigned long long)c->size);
    int mode = ids ? 1 : 2;
    UCFP_REQUIRE(c->id_mode == 0 || c->id_mode == mode, UCFP_E_STATE, "corpus mixes explicit and implicit record ids");
    cudaStream_t st = c->ctx->stream;
    size_t rb = row_bytes(c);
    if (mode == 1 && !c->ids) UCFP_CUDA_TRY(cudaMalloc((void **)&c->ids, 8 * (c->capacity + 16)));
    cudaMemcpyKind kr = classify(rows) == Mem::Device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    UCFP_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(c->rows) + rb * c->size, rows, rb * n, kr, st));
    if (mode == 1) {
        cudaMemcpyKind ki = classify(ids) == Mem::Device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        UCFP_CUDA_TRY(cudaMemcpyAsync(c->ids + c->size, ids, 8 * n, ki, st));
    }
    UCFP_TRY(after_append(c, c->size, n));
    // host sources may be reused by the caller as soon as we return
    if (kr == cudaMemcpyHostToDevice || mode == 1) UCFP_CUDA_TRY(cudaStreamSynchronize(st));
    c->id_mode = mode;
    c->size += n;
    return UCFP_OK;
}

int ucfp_corpus_append_strided(ucfp_corpus *c, const uint64_t *ids, const void *records, uint64_t record_stride,
                               uint64_t field_offset, uint64_t n) {
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    UCFP_GUARD(c->ctx);
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(records != nullptr, UCFP_E_INVALID, "records is NULL");
    size_t rb = row_bytes(c);
    UCFP_REQUIRE(record_stride >= field_offset + rb, UCFP_E_INVALID, "record stride %llu cannot hold a %zu-byte field at offset %llu",
                 (unsigned long long)record_stride, rb, (unsigned long long)field_offset);
    UCFP_REQUIRE(c->size + n <= c->capacity, UCFP_E_CAPACITY, "append of %llu rows exceeds capacity %llu (size %llu)",
                 (unsigned long long)n, (unsigned long long)c->capacity, (unsigned long long)c->size);
    int mode = ids ? 1 : 2;
    UCFP_REQUIRE(c->id_mode == 0 || c->id_mode == mode, UCFP_E_STATE, "corpus mixes explicit and implicit record ids");
    cudaStream_t st = c->ctx->stream;
    if (mode == 1 && !c->ids) UCFP_CUDA_TRY(cudaMalloc((void **)&c->ids, 8 * (c->capacity + 16)));
    const bool dev_src = classify(records) == Mem::Device;
    // a pitched copy gathers the field of every record: source pitch = record stride, width = one row
    UCFP_CUDA_TRY(cudaMemcpy2DAsync(static_cast<char *>(c->rows) + rb * c->size, rb, static_cast<const char *>(records) + field_offset,
                                    record_stride, rb, n, dev_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (mode == 1) {
        cudaMemcpyKind ki = classify(ids) == Mem::Device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        UCFP_CUDA_TRY(cudaMemcpyAsync(c->ids + c->size, ids, 8 * n, ki, st));
    }
    UCFP_TRY(after_append(c, c->size, n));
    if (!dev_src || mode == 1) UCFP_CUDA_TRY(cudaStreamSynchronize(st));
    c->id_mode = mode;
    c->size += n;
    return UCFP_OK;
}

int ucfp_corpus_append_synthetic(ucfp_corpus *c, uint64_t seed, uint64_t start_row, uint64_t n) {
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    UCFP_GUARD(c->ctx);
    UCFP_REQUIRE(c->kind == UCFP_KIND_HAMMING64 || c->kind == UCFP_KIND_MINHASH128, UCFP_E_UNSUPPORTED,
                 "synthetic rows exist for HAMMING64 and MINHASH128 corpora only");
    UCFP_REQUIRE(c->size + n <= c->capacity, UCFP_E_CAPACITY, "append of %llu rows exceeds capacity", (unsigned long long)n);
    UCFP_REQUIRE(c->id_mode == 0 || c->id_mode == 2, UCFP_E_STATE, "corpus mixes explicit and implicit record ids");
    uint64_t wpr = c->kind == UCFP_KIND_HAMMING64 ? 1 : 128;
    UCFP_TRY(synth_fill_u64(c->ctx, static_cast<uint64_t *>(c->rows) + c->size * wpr, n * wpr, seed, start_row * wpr));
    UCFP_TRY(after_append(c, c->size, n));
    c->id_mode = 2;
    c->size += n;
    return UCFP_OK;
}

// ---- scans --------------------------------------------------------------------------------

}  // extern "C"

template <typename Key, typename ScanFn>
static int run_scan(ucfp_corpus *c, int want_kind, const void *queries, size_t q_bytes, size_t nq, size_t k, uint64_t *ids_out,
                    Key *keys_out, ScanFn scan) {
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    ucfp_ctx *ctx = c->ctx;
    UCFP_GUARD(ctx);
    UCFP_REQUIRE(c->kind == want_kind, UCFP_E_STATE, "corpus kind %d cannot serve this scan (needs kind %d)", c->kind, want_kind);
    if (nq == 0 || k == 0) return UCFP_OK;
    UCFP_REQUIRE(queries && ids_out && keys_out, UCFP_E_INVALID, "NULL query or output buffer");
    const void *q_dev = nullptr;
    UCFP_TRY(stage_in(ctx, ctx->q_dev, queries, q_bytes, &q_dev));
    void *ids_dev = nullptr, *keys_dev = nullptr;
    bool ids_host = false, keys_host = false;
    UCFP_TRY(stage_out(ctx->out_ids_dev, ids_out, 8 * nq * k, &ids_dev, &ids_host));
    UCFP_TRY(stage_out(ctx->out_keys_dev, keys_out, sizeof(Key) * nq * k, &keys_dev, &keys_host));
    UCFP_TRY(stats_reset(ctx));
    UCFP_TRY(scan(q_dev, static_cast<uint64_t *>(ids_dev), static_cast<Key *>(keys_dev)));
    if (ids_host) UCFP_TRY(copy_back(ctx, ids_out, ids_dev, 8 * nq * k));
    if (keys_host) UCFP_TRY(copy_back(ctx, keys_out, keys_dev, sizeof(Key) * nq * k));
    if (ids_host || keys_host) UCFP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return UCFP_OK;
}

extern "C" {

int ucfp_scan_hamming(ucfp_corpus *c, const uint64_t *queries, size_t nq, size_t k, uint64_t *ids_out, uint32_t *dist_out) {
    return run_scan<uint32_t>(c, UCFP_KIND_HAMMING64, queries, 8 * nq, nq, k, ids_out, dist_out,
                              [&](const void *q, uint64_t *io, uint32_t *ko) {
                                  return hamming_scan(c, static_cast<const uint64_t *>(q), nq, k, io, ko);
                              });
}

int ucfp_scan_jaccard(ucfp_corpus *c, const uint64_t *queries, size_t nq, size_t k, uint64_t *ids_out, uint32_t *matches_out) {
    return run_scan<uint32_t>(c, UCFP_KIND_MINHASH128, queries, 1024 * nq, nq, k, ids_out, matches_out,
                              [&](const void *q, uint64_t *io, uint32_t *ko) {
                                  return jaccard_scan(c, static_cast<const uint64_t *>(q), nq, k, io, ko);
                              });
}

int ucfp_scan_cosine(ucfp_corpus *c, const float *queries, size_t nq, size_t k, uint64_t *ids_out, float *score_out) {
    size_t dim = c ? c->dim : 0;
    return run_scan<float>(c, UCFP_KIND_COSINE, queries, 4 * dim * nq, nq, k, ids_out, score_out,
                           [&](const void *q, uint64_t *io, float *ko) {
                               return cosine_scan(c, static_cast<const float *>(q), nq, k, io, ko);
                           });
}

}  // extern "C"

extern "C" int ucfp_ctx_last_scan_fallbacks(ucfp_ctx *ctx, uint64_t *queries_recomputed) {
    UCFP_GUARD(ctx);
    UCFP_REQUIRE(queries_recomputed != nullptr, UCFP_E_INVALID, "NULL output");
    *queries_recomputed = 0;
    if (!ctx->stats.ptr) return UCFP_OK;
    UCFP_CUDA_TRY(cudaMemcpyAsync(queries_recomputed, ctx->stats.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream));
    UCFP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return UCFP_OK;
}

template <typename Key, typename MergeFn>
static int run_merge(ucfp_ctx *ctx, const uint64_t *ids_in, const Key *keys_in, size_t parts, size_t nq, size_t k,
                     uint64_t *ids_out, Key *keys_out, MergeFn merge) {
    UCFP_GUARD(ctx);
    if (nq == 0 || k == 0 || parts == 0) return UCFP_OK;
    UCFP_REQUIRE(ids_in && keys_in && ids_out && keys_out, UCFP_E_INVALID, "NULL buffer");
    size_t n_in = parts * nq * k;
    const void *ids_in_dev = nullptr, *keys_in_dev = nullptr;
    UCFP_TRY(stage_in(ctx, ctx->cand, ids_in, 8 * n_in, &ids_in_dev));
    UCFP_TRY(stage_in(ctx, ctx->misc, keys_in, sizeof(Key) * n_in, &keys_in_dev));
    void *ids_dev = nullptr, *keys_dev = nullptr;
    bool ids_host = false, keys_host = false;
    UCFP_TRY(stage_out(ctx->out_ids_dev, ids_out, 8 * nq * k, &ids_dev, &ids_host));
    UCFP_TRY(stage_out(ctx->out_keys_dev, keys_out, sizeof(Key) * nq * k, &keys_dev, &keys_host));
    UCFP_TRY(merge(static_cast<const uint64_t *>(ids_in_dev), static_cast<const Key *>(keys_in_dev),
                   static_cast<uint64_t *>(ids_dev), static_cast<Key *>(keys_dev)));
    if (ids_host) UCFP_TRY(copy_back(ctx, ids_out, ids_dev, 8 * nq * k));
    if (keys_host) UCFP_TRY(copy_back(ctx, keys_out, keys_dev, sizeof(Key) * nq * k));
    if (ids_host || keys_host) UCFP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return UCFP_OK;
}

extern "C" {

int ucfp_merge_topk_u32(ucfp_ctx *ctx, const uint64_t *ids_in, const uint32_t *keys_in, size_t parts, size_t nq, size_t k,
                        int descending, uint64_t *ids_out, uint32_t *keys_out) {
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    return run_merge<uint32_t>(ctx, ids_in, keys_in, parts, nq, k, ids_out, keys_out,
                               [&](const uint64_t *ii, const uint32_t *ki, uint64_t *io, uint32_t *ko) {
                                   return merge_u32(ctx, ii, ki, parts, nq, k, descending, io, ko);
                               });
}

int ucfp_merge_topk_f32(ucfp_ctx *ctx, const uint64_t *ids_in, const float *scores_in, size_t parts, size_t nq, size_t k,
                        uint64_t *ids_out, float *scores_out) {
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    return run_merge<float>(ctx, ids_in, scores_in, parts, nq, k, ids_out, scores_out,
                            [&](const uint64_t *ii, const float *ki, uint64_t *io, float *ko) {
                                return merge_f32(ctx, ii, ki, parts, nq, k, io, ko);
                            });
}

// ---- image hashing ------------------------------------------------------------------------

int ucfp_image_hash_batch(ucfp_ctx *ctx, const ucfp_image_desc *imgs, size_t n, uint32_t algo_mask, ucfp_image_hashes *out,
                          int32_t *status) {
    UCFP_GUARD(ctx);
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(imgs && out, UCFP_E_INVALID, "NULL image descriptors or output");
    UCFP_REQUIRE((algo_mask & ~UCFP_ALGO_MULTI) == 0 && algo_mask != 0, UCFP_E_INVALID, "bad algo_mask 0x%x", algo_mask);
    void *out_dev = nullptr;
    bool out_host = false;
    UCFP_TRY(stage_out(ctx->img_out_dev, out, sizeof(ucfp_image_hashes) * n, &out_dev, &out_host));
    std::vector<int32_t> st_host(n, 0);
    UCFP_TRY(image_hash_batch(ctx, imgs, n, algo_mask, static_cast<ucfp_image_hashes *>(out_dev), st_host.data()));
    if (out_host) UCFP_TRY(copy_back(ctx, out, out_dev, sizeof(ucfp_image_hashes) * n));
    if (status) {
        if (classify(status) == Mem::Device)
            UCFP_CUDA_TRY(cudaMemcpyAsync(status, st_host.data(), 4 * n, cudaMemcpyHostToDevice, ctx->stream));
        else
            memcpy(status, st_host.data(), 4 * n);
    }
    if (out_host || status) UCFP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return UCFP_OK;
}

int ucfp_image_hash_uniform(ucfp_ctx *ctx, const uint8_t *pixels, size_t n, uint32_t width, uint32_t height, uint64_t row_stride,
                            uint64_t image_stride, uint32_t algo_mask, ucfp_image_hashes *out) {
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(pixels && out, UCFP_E_INVALID, "NULL pixels or output");
    UCFP_REQUIRE(width >= 4 && height >= 4 && row_stride >= 3ull * width && image_stride >= row_stride * (height - 1) + 3ull * width,
                 UCFP_E_INVALID, "bad image geometry %ux%u stride %llu/%llu", width, height, (unsigned long long)row_stride,
                 (unsigned long long)image_stride);
    std::vector<ucfp_image_desc> d(n);
    for (size_t i = 0; i < n; ++i) d[i] = ucfp_image_desc{pixels + i * image_stride, width, height, row_stride};
    std::vector<int32_t> st(n, 0);
    int rc = ucfp_image_hash_batch(ctx, d.data(), n, algo_mask, out, st.data());
    if (rc != UCFP_OK) return rc;
    for (size_t i = 0; i < n; ++i)
        if (st[i] != UCFP_OK) { set_error("image %zu failed with status %d", i, st[i]); return st[i]; }
    return UCFP_OK;
}

}  // extern "C"
