// group.cu -- one query batch against a corpus sharded by record range over the GPUs of one box.
//
// Process models: (a) ONE process drives n GPUs (the FFI use-case: the Rust host) -- ucfp_group_create makes a context
// per device and an NCCL communicator over them with ncclCommInitAll; a worker thread per device enqueues that
// device's part so that the n shards are launched concurrently; (b) one process per GPU (the bench harness under
// torchrun) -- ucfp_group_unique_id / ucfp_group_join build the communicator with ncclCommInitRank.
//
// Data path of a rank: stage the query batch, scan the shard (the Hamming scan folds the other ranks' admission bounds in
// between chunks, topk_select.cuh: exchange_bounds), pack the local top-k as 16-byte (id, key) records, ONE ncclAllGather
// of Q x k records, deterministic merge of the G x k candidates per query (merge.cu) on the rank that owns the output.
// NCCL is resolved with dlopen at the first group call: the library has no load-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; the symbols are looked up at run time

#include <functional>
#include <thread>

#include "api_util.cuh"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

std::mutex g_nccl_mu;
NcclApi g_nccl;

int nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return UCFP_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    UCFP_REQUIRE(h != nullptr, UCFP_E_UNSUPPORTED, "NCCL is not available: %s", dlerror());
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    UCFP_REQUIRE(a.GetUniqueId && a.CommInitAll && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString, UCFP_E_UNSUPPORTED,
                 "libnccl lacks a required symbol");
    g_nccl = a;
    return UCFP_OK;
}

#define UCFP_NCCL_TRY(expr)                                                                          \
    do {                                                                                             \
        ncclResult_t _r = (expr);                                                                    \
        if (_r != ncclSuccess) {                                                                     \
            ::ucfp::set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r));                    \
            return UCFP_E_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

// How many of a query pass's bound exchanges (topk_select.cuh) really run.  Measured on 8 B200s, 1 B Hamming codes, 1 024 queries
// (profiles/r02_bound_exchange_n8.md): 0 / 1 / 2 / 4 exchanges = 205.3 / 204.3 / 203.1 / 201.3 K queries/s with the scan kernels at
// 4.70 ms in every case -- since fired strips are parked (hamming.cu) a looser bound costs a shard nothing measurable, and each
// exchange is ~25 us of pack + all-gather + fold.  The Jaccard scan did not gain either (33.0 K vs 33.4 K queries/s).  Default 0;
// UCFP_GROUP_EXCHANGES=n (read when the group is created, and it must agree on every rank) turns them on for corpora whose shards
// are very unequal in how close their rows are to the queries.
int bound_exchanges() {
    const char *e = getenv("UCFP_GROUP_EXCHANGES");
    const long n = e ? atol(e) : 0;
    return (int)(n < 0 ? 0 : n > 64 ? 64 : n);
}

int allgather_hook(void *comm, const void *send, void *recv, size_t bytes, cudaStream_t st) {
    UCFP_NCCL_TRY(g_nccl.AllGather(send, recv, bytes, ncclChar, static_cast<ncclComm_t>(comm), st));
    return UCFP_OK;
}

// A persistent host thread per local rank (single-process groups of more than one GPU): the shards' launch sequences
// (~30 launches each) are enqueued side by side instead of one device after the other.
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, stop = false;
    int rc = UCFP_OK;
    char err[256] = "";
    void run() {
        for (;;) {
            std::function<int()> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return has_job || stop; });
                if (stop) return;
                j = std::move(job);
                has_job = false;
            }
            int r;
            try { r = j(); } catch (...) { ucfp::set_error("internal error in a group worker"); r = UCFP_E_STATE; }
            {
                std::lock_guard<std::mutex> lk(mu);
                rc = r;
                snprintf(err, sizeof(err), "%s", r == UCFP_OK ? "" : ucfp_last_error());
                done = true;
            }
            cv.notify_all();
        }
    }
};

struct Rank {
    ucfp_ctx *ctx = nullptr;
    bool own_ctx = false;
    ncclComm_t comm = nullptr;
    int world_rank = 0;
    ucfp_exchange xch;              // bound-exchange hook and its buffers
    ucfp::DevBuf send, recv;        // packed top-k records: local list, gathered lists
    Worker *worker = nullptr;
};

}  // namespace

struct ucfp_group {
    int n_local = 0, world = 0;
    std::vector<Rank> ranks;
    std::mutex mu;   // one group scan at a time (collectives of concurrent scans must not interleave)
};

namespace {

struct ScanCall {
    int kind; const void *queries; size_t nq, k; uint64_t *ids_out; void *keys_out;
    int q_device;     // device that holds `queries` (-1: host)
    int out_device;   // device that holds the outputs (-1: host)
};

int pointer_device(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}

// The part of a group scan one rank runs (its device is made current by the lane lease).
int rank_scan(ucfp_group *g, int local, ucfp_corpus *c, const ScanCall &s) {
    using namespace ucfp;
    Rank &R = g->ranks[local];
    UCFP_REQUIRE(c != nullptr && c->ctx == R.ctx, UCFP_E_INVALID, "corpora[%d] does not live on the context of local rank %d", local, local);
    UCFP_REQUIRE(c->kind == s.kind, UCFP_E_STATE, "corpus kind %d cannot serve this scan (needs kind %d)", c->kind, s.kind);
    std::shared_lock<std::shared_mutex> rl(c->rw);
    UCFP_LEASE(R.ctx);
    const size_t q_bytes = s.nq * row_bytes(c), n_out = s.nq * s.k;
    // queries: host -> H2D; this device -> in place; another local device -> peer copy
    const void *q_dev = nullptr;
    if (s.q_device == R.ctx->device) q_dev = s.queries;
    else {
        UCFP_TRY(lane->q_dev.reserve(q_bytes));
        if (s.q_device < 0) UCFP_CUDA_TRY(cudaMemcpyAsync(lane->q_dev.ptr, s.queries, q_bytes, cudaMemcpyHostToDevice, lane->stream));
        else UCFP_CUDA_TRY(cudaMemcpyPeerAsync(lane->q_dev.ptr, R.ctx->device, s.queries, s.q_device, q_bytes, lane->stream));
        q_dev = lane->q_dev.ptr;
    }
    UCFP_TRY(lane->out_ids_dev.reserve(8 * n_out));
    UCFP_TRY(lane->out_keys_dev.reserve(4 * n_out));
    uint64_t *ids_loc = lane->out_ids_dev.as<uint64_t>();
    void *keys_loc = lane->out_keys_dev.ptr;
    UCFP_TRY(stats_reset(lane));
    R.xch.done = 0;
    lane->xch = g->world > 1 ? &R.xch : nullptr;
    int rc;
    if (s.kind == UCFP_KIND_HAMMING64) rc = hamming_scan(lane, c, static_cast<const uint64_t *>(q_dev), s.nq, s.k, ids_loc, static_cast<uint32_t *>(keys_loc));
    else if (s.kind == UCFP_KIND_MINHASH128) rc = jaccard_scan(lane, c, static_cast<const uint64_t *>(q_dev), s.nq, s.k, ids_loc, static_cast<uint32_t *>(keys_loc));
    else rc = cosine_scan(lane, c, static_cast<const float *>(q_dev), s.nq, s.k, ids_loc, static_cast<float *>(keys_loc));
    lane->xch = nullptr;
    UCFP_TRY(rc);
    // the rank that owns the output memory merges; with host outputs that is local rank 0
    const bool owner = s.out_device < 0 ? local == 0 : s.out_device == R.ctx->device;
    uint64_t *ids_final = ids_loc;
    void *keys_final = keys_loc;
    if (g->world > 1) {
        UCFP_TRY(R.send.reserve(16 * n_out));
        UCFP_TRY(R.recv.reserve(16 * n_out * (size_t)g->world));
        UCFP_TRY(pack_topk(lane, ids_loc, keys_loc, n_out, R.send.ptr));
        UCFP_NCCL_TRY(g_nccl.AllGather(R.send.ptr, R.recv.ptr, 16 * n_out, ncclChar, R.comm, lane->stream));
        if (owner) {
            if (s.out_device >= 0) { ids_final = s.ids_out; keys_final = s.keys_out; }   // straight into the caller's device buffers
            else {
                UCFP_TRY(lane->cand.reserve(12 * n_out));
                ids_final = lane->cand.as<uint64_t>();
                keys_final = ids_final + n_out;
            }
            UCFP_TRY(merge_packed(lane, R.recv.ptr, (size_t)g->world, s.nq, s.k, s.kind == UCFP_KIND_COSINE, s.kind == UCFP_KIND_MINHASH128 ? 1 : 0,
                                  ids_final, keys_final));
        }
    } else if (owner && s.out_device >= 0) {
        UCFP_CUDA_TRY(cudaMemcpyAsync(s.ids_out, ids_loc, 8 * n_out, cudaMemcpyDeviceToDevice, lane->stream));
        UCFP_CUDA_TRY(cudaMemcpyAsync(s.keys_out, keys_loc, 4 * n_out, cudaMemcpyDeviceToDevice, lane->stream));
    }
    if (owner && s.out_device < 0) {
        UCFP_TRY(copy_back(lane, s.ids_out, ids_final, 8 * n_out));
        UCFP_TRY(copy_back(lane, s.keys_out, keys_final, 4 * n_out));
    }
    {
        std::lock_guard<std::mutex> lk(R.ctx->mu);
        for (int i = 0; i < R.ctx->n_lanes; ++i) if (R.ctx->lanes[i] == lane) R.ctx->last_scan_lane = i;
    }
    // Pooled mode: complete before returning (the rank's send / receive buffers would otherwise be reused by the next scan
    // on another lane's stream).  Shared-stream mode: consecutive scans are ordered on the caller's stream, so a scan with
    // device outputs returns asynchronously like every other entry point.
    return finish_call(lane, owner && s.out_device < 0);
}

int group_scan(ucfp_group *g, ucfp_corpus *const *corpora, int kind, const void *queries, size_t nq, size_t k, uint64_t *ids_out, void *keys_out) {
    UCFP_REQUIRE(g != nullptr, UCFP_E_INVALID, "null group");
    UCFP_REQUIRE(corpora != nullptr, UCFP_E_INVALID, "corpora is NULL");
    if (nq == 0 || k == 0) return UCFP_OK;
    UCFP_REQUIRE(queries && ids_out && keys_out, UCFP_E_INVALID, "NULL query or output buffer");
    UCFP_REQUIRE((size_t)g->world * k <= 16384, UCFP_E_UNSUPPORTED, "group merge supports world * k <= 16384");
    std::lock_guard<std::mutex> lk(g->mu);
    ScanCall s{kind, queries, nq, k, ids_out, keys_out, pointer_device(queries), pointer_device(ids_out)};
    UCFP_REQUIRE(pointer_device(keys_out) == s.out_device, UCFP_E_INVALID, "ids_out and keys_out must live in the same memory");
    if (s.out_device >= 0) {
        bool local = false;
        for (Rank &R : g->ranks) local = local || R.ctx->device == s.out_device;
        UCFP_REQUIRE(local, UCFP_E_INVALID, "the output buffers live on device %d, which this group does not drive", s.out_device);
    }
    if (g->n_local == 1) return rank_scan(g, 0, corpora[0], s);
    for (int i = 0; i < g->n_local; ++i) {
        Worker *w = g->ranks[i].worker;
        ucfp_corpus *c = corpora[i];
        std::lock_guard<std::mutex> wl(w->mu);
        w->job = [g, i, c, s]() { return rank_scan(g, i, c, s); };
        w->has_job = true; w->done = false;
        w->cv.notify_all();
    }
    int rc = UCFP_OK;
    for (int i = 0; i < g->n_local; ++i) {
        Worker *w = g->ranks[i].worker;
        std::unique_lock<std::mutex> wl(w->mu);
        w->cv.wait(wl, [&] { return w->done; });
        if (w->rc != UCFP_OK && rc == UCFP_OK) { rc = w->rc; ucfp::set_error("local rank %d: %s", i, w->err); }
    }
    return rc;
}

void group_free(ucfp_group *g) {
    for (Rank &R : g->ranks) {
        if (R.worker) {
            { std::lock_guard<std::mutex> lk(R.worker->mu); R.worker->stop = true; }
            R.worker->cv.notify_all();
            if (R.worker->th.joinable()) R.worker->th.join();
            delete R.worker;
        }
        if (R.ctx) {
            ucfp::DeviceGuard dg(R.ctx->device);
            ucfp_ctx_synchronize(R.ctx);
            if (R.comm) g_nccl.CommDestroy(R.comm);
            R.send.release(); R.recv.release(); R.xch.send.release(); R.xch.recv.release();
            if (R.own_ctx) ucfp_destroy(R.ctx);
        }
    }
    delete g;
}

}  // namespace

extern "C" {

int ucfp_group_create(const int *devices, int n, ucfp_group **out) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(out != nullptr, UCFP_E_INVALID, "ucfp_group_create: out is NULL");
    *out = nullptr;
    UCFP_REQUIRE(devices != nullptr && n >= 1 && n <= 64, UCFP_E_INVALID, "bad device list");
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) UCFP_REQUIRE(devices[i] != devices[j], UCFP_E_INVALID, "device %d is listed twice", devices[i]);
    if (n > 1) UCFP_TRY(nccl_load());
    ucfp_group *g = new ucfp_group();
    g->n_local = n; g->world = n;
    g->ranks.resize(n);
    for (int i = 0; i < n; ++i) {
        int rc = ucfp_init(devices[i], &g->ranks[i].ctx);
        if (rc != UCFP_OK) { group_free(g); return rc; }
        g->ranks[i].own_ctx = true;
        g->ranks[i].world_rank = i;
    }
    if (n > 1) {
        std::vector<ncclComm_t> comms(n);
        ncclResult_t r = g_nccl.CommInitAll(comms.data(), n, devices);
        if (r != ncclSuccess) { ucfp::set_error("ncclCommInitAll failed: %s", g_nccl.GetErrorString(r)); group_free(g); return UCFP_E_CUDA; }
        for (int i = 0; i < n; ++i) {
            Rank &R = g->ranks[i];
            R.comm = comms[i];
            R.xch.comm = comms[i]; R.xch.world = n; R.xch.allgather = allgather_hook; R.xch.max_real = bound_exchanges();
            R.worker = new Worker();
            R.worker->th = std::thread(&Worker::run, R.worker);
        }
    }
    *out = g;
    return UCFP_OK;
    UCFP_API_END
}

int ucfp_group_unique_id(void *id128) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(id128 != nullptr, UCFP_E_INVALID, "id128 is NULL");
    UCFP_TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "the ABI ships NCCL unique ids as 128 bytes");
    ncclUniqueId id;
    UCFP_NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return UCFP_OK;
    UCFP_API_END
}

int ucfp_group_join(ucfp_ctx *ctx, const void *id128, int rank, int world, ucfp_group **out) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(out != nullptr, UCFP_E_INVALID, "ucfp_group_join: out is NULL");
    *out = nullptr;
    UCFP_REQUIRE(ctx != nullptr && id128 != nullptr, UCFP_E_INVALID, "null context or id");
    UCFP_REQUIRE(world >= 1 && rank >= 0 && rank < world, UCFP_E_INVALID, "bad rank %d of %d", rank, world);
    if (world > 1) UCFP_TRY(nccl_load());
    ucfp_group *g = new ucfp_group();
    g->n_local = 1; g->world = world;
    g->ranks.resize(1);
    Rank &R = g->ranks[0];
    R.ctx = ctx; R.own_ctx = false; R.world_rank = rank;
    if (world > 1) {
        ucfp::DeviceGuard dg(ctx->device);
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclResult_t r = g_nccl.CommInitRank(&R.comm, world, id, rank);
        if (r != ncclSuccess) { ucfp::set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); R.ctx = nullptr; group_free(g); return UCFP_E_CUDA; }
        R.xch.comm = R.comm; R.xch.world = world; R.xch.allgather = allgather_hook; R.xch.max_real = bound_exchanges();
    }
    *out = g;
    return UCFP_OK;
    UCFP_API_END
}

void ucfp_group_destroy(ucfp_group *g) {
    if (!g) return;
    try { group_free(g); } catch (...) {}
}

int ucfp_group_local_size(const ucfp_group *g) { return g ? g->n_local : 0; }
int ucfp_group_world_size(const ucfp_group *g) { return g ? g->world : 0; }
ucfp_ctx *ucfp_group_ctx(ucfp_group *g, int local_rank) { return (g && local_rank >= 0 && local_rank < g->n_local) ? g->ranks[local_rank].ctx : nullptr; }

int ucfp_group_scan_hamming(ucfp_group *g, ucfp_corpus *const *corpora, const uint64_t *queries, size_t nq, size_t k, uint64_t *ids_out,
                            uint32_t *dist_out) {
    UCFP_API_BEGIN
    return group_scan(g, corpora, UCFP_KIND_HAMMING64, queries, nq, k, ids_out, dist_out);
    UCFP_API_END
}

int ucfp_group_scan_jaccard(ucfp_group *g, ucfp_corpus *const *corpora, const uint64_t *queries, size_t nq, size_t k, uint64_t *ids_out,
                            uint32_t *matches_out) {
    UCFP_API_BEGIN
    return group_scan(g, corpora, UCFP_KIND_MINHASH128, queries, nq, k, ids_out, matches_out);
    UCFP_API_END
}

int ucfp_group_scan_cosine(ucfp_group *g, ucfp_corpus *const *corpora, const float *queries, size_t nq, size_t k, uint64_t *ids_out,
                           float *score_out) {
    UCFP_API_BEGIN
    return group_scan(g, corpora, UCFP_KIND_COSINE, queries, nq, k, ids_out, score_out);
    UCFP_API_END
}

}  // extern "C"
