// api_util.cuh -- what the files that implement extern "C" entry points share: the exception barrier, lane leases,
// host/device staging helpers.  (api.cu, corpus.cu, batcher.cu, group.cu, nvjpeg.cu)
#pragma once

#include "common.cuh"

// Nothing may unwind across the C ABI (the reference's release profile is panic = "abort", Cargo.toml:180): every
// entry-point body sits between these two.
#define UCFP_API_BEGIN try {
#define UCFP_API_END                                                                              \
    }                                                                                             \
    catch (const std::bad_alloc &) {                                                              \
        ::ucfp::set_error("out of host memory");                                                  \
        return UCFP_E_OOM;                                                                        \
    }                                                                                             \
    catch (const std::exception &e) {                                                             \
        ::ucfp::set_error("internal error: %s", e.what());                                        \
        return UCFP_E_STATE;                                                                      \
    }                                                                                             \
    catch (...) {                                                                                 \
        ::ucfp::set_error("internal error");                                                      \
        return UCFP_E_STATE;                                                                      \
    }

namespace ucfp {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) { ok = false; cudaGetLastError(); }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Owns one lane of the context (and makes the context's device current) until it goes out of scope.  Pooled mode: a
// free lane, a new one while fewer than kUcfpMaxLanes exist, else waits.  Shared-stream mode: lane 0, one call at a time.
struct LaneLease {
    ucfp_ctx *ctx;
    ucfp_lane *lane = nullptr;
    int prev_device = -1;
    explicit LaneLease(ucfp_ctx *c);
    ~LaneLease();
    LaneLease(const LaneLease &) = delete;
    LaneLease &operator=(const LaneLease &) = delete;
};

#define UCFP_LEASE(ctxp)                                              \
    ::ucfp::LaneLease _lease(ctxp);                                   \
    if (!_lease.lane) return UCFP_E_CUDA;                             \
    ucfp_lane *lane = _lease.lane

int stage_in(ucfp_lane *ln, DevBuf &buf, const void *user, size_t bytes, const void **dev);
int stage_out(DevBuf &buf, void *user, size_t bytes, void **dev, bool *is_host);
int copy_back(ucfp_lane *ln, void *user, const void *dev, size_t bytes);
int finish_call(ucfp_lane *ln, bool host_outputs);
size_t row_bytes(const ucfp_corpus *c);
int after_append(ucfp_lane *ln, ucfp_corpus *c, uint64_t first, uint64_t n);
int run_scan_any(ucfp_corpus *c, int want_kind, const void *queries, size_t nq, size_t k, uint64_t *ids_out, void *keys_out);

// corpus.cu: allocation, growth, delete, insert-or-replace
int corpus_device_init(ucfp_ctx *ctx);
int corpus_alloc_arrays(ucfp_lane *ln, ucfp_corpus *c, uint64_t capacity);
int corpus_grow(ucfp_lane *ln, ucfp_corpus *c, uint64_t capacity);
int corpus_delete_ids(ucfp_lane *ln, ucfp_corpus *c, const uint64_t *ids, uint64_t n, uint64_t *n_removed);
int corpus_upsert_rows(ucfp_lane *ln, ucfp_corpus *c, const uint64_t *ids, const void *rows, uint64_t n, uint64_t *n_replaced);

}  // namespace ucfp
