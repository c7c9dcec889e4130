// batcher.cu -- coalesces the single-query calls of many host threads into batched scans (include/ucfp_cuda.h,
// "query batcher").  The reference answers one query per request with up to 512 requests in flight
// (src/bin/ucfp.rs:262-267; handlers::query, src/server/handlers.rs:143-187); a scan call per query would pay one full
// HBM pass per query.  Host-side only: the scans themselves are the ordinary entry points (run_scan_any).
#include <chrono>
#include <deque>
#include <thread>

#include "api_util.cuh"

namespace {

struct Request {
    const void *query;
    size_t k;
    uint64_t *ids_out;
    void *keys_out;
    int rc = UCFP_OK;
    bool done = false;
    char err[160] = "";
    std::chrono::steady_clock::time_point t_in;
};

}  // namespace

struct ucfp_batcher {
    ucfp_corpus *corpus = nullptr;
    int kind = 0;
    size_t q_bytes = 0;
    uint32_t max_batch = 0, max_delay_us = 0;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Request *> queue;
    bool stop = false;
    std::vector<std::thread> workers;
    std::atomic<uint64_t> n_queries{0}, n_batches{0}, largest{0};
};

namespace {

constexpr int kWorkers = 2;   // one batch scans while the next is assembled and the previous one's results are handed out

void worker_main(ucfp_batcher *b) {
    std::vector<unsigned char> qbuf;
    std::vector<uint64_t> ids;
    std::vector<uint32_t> keys;
    std::vector<Request *> batch;
    for (;;) {
        batch.clear();
        {
            std::unique_lock<std::mutex> lk(b->mu);
            b->cv_work.wait(lk, [&] { return b->stop || !b->queue.empty(); });
            if (b->stop && b->queue.empty()) return;
            // the first query in line waits at most max_delay_us for company; a full batch leaves at once
            const auto deadline = b->queue.front()->t_in + std::chrono::microseconds(b->max_delay_us);
            while (!b->stop && b->queue.size() < b->max_batch && std::chrono::steady_clock::now() < deadline)
                b->cv_work.wait_until(lk, deadline);
            while (!b->queue.empty() && batch.size() < b->max_batch) { batch.push_back(b->queue.front()); b->queue.pop_front(); }
        }
        if (batch.empty()) continue;
        const size_t nq = batch.size();
        size_t k = 0;
        for (Request *r : batch) k = r->k > k ? r->k : k;
        int rc = UCFP_OK;
        try {
            qbuf.resize(nq * b->q_bytes);
            ids.resize(nq * k);
            keys.resize(nq * k);
            for (size_t i = 0; i < nq; ++i) memcpy(qbuf.data() + i * b->q_bytes, batch[i]->query, b->q_bytes);
            rc = ucfp::run_scan_any(b->corpus, b->kind, qbuf.data(), nq, k, ids.data(), keys.data());
        } catch (const std::bad_alloc &) {
            ucfp::set_error("out of host memory");
            rc = UCFP_E_OOM;
        } catch (...) {
            ucfp::set_error("internal error in the batcher");
            rc = UCFP_E_STATE;
        }
        const char *msg = rc == UCFP_OK ? "" : ucfp_last_error();   // this worker thread's message: handed to every caller of the batch
        for (size_t i = 0; i < nq; ++i) {
            Request *r = batch[i];
            if (rc == UCFP_OK) {
                memcpy(r->ids_out, ids.data() + i * k, 8 * r->k);
                memcpy(r->keys_out, keys.data() + i * k, 4 * r->k);
            } else {
                snprintf(r->err, sizeof(r->err), "%s", msg);
            }
            r->rc = rc;
        }
        b->n_queries.fetch_add(nq);
        b->n_batches.fetch_add(1);
        uint64_t seen = b->largest.load();
        while (nq > seen && !b->largest.compare_exchange_weak(seen, nq)) {}
        {
            std::lock_guard<std::mutex> lk(b->mu);
            for (Request *r : batch) r->done = true;
        }
        b->cv_done.notify_all();
    }
}

}  // namespace

extern "C" {

int ucfp_batcher_create(ucfp_corpus *c, uint32_t max_batch, uint32_t max_delay_us, ucfp_batcher **out) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(out != nullptr, UCFP_E_INVALID, "ucfp_batcher_create: out is NULL");
    *out = nullptr;
    UCFP_REQUIRE(c != nullptr, UCFP_E_INVALID, "null corpus");
    UCFP_REQUIRE(max_batch >= 1 && max_batch <= 4096, UCFP_E_INVALID, "max_batch must be in 1..4096 (got %u)", max_batch);
    ucfp_batcher *b = new ucfp_batcher();
    b->corpus = c; b->kind = c->kind; b->q_bytes = ucfp::row_bytes(c);
    b->max_batch = max_batch; b->max_delay_us = max_delay_us;
    try {
        for (int i = 0; i < kWorkers; ++i) b->workers.emplace_back(worker_main, b);
    } catch (...) {
        { std::lock_guard<std::mutex> lk(b->mu); b->stop = true; }
        b->cv_work.notify_all();
        for (auto &t : b->workers) t.join();
        delete b;
        ucfp::set_error("cannot start the batcher's worker threads");
        return UCFP_E_STATE;
    }
    *out = b;
    return UCFP_OK;
    UCFP_API_END
}

void ucfp_batcher_destroy(ucfp_batcher *b) {
    if (!b) return;
    try {
        { std::lock_guard<std::mutex> lk(b->mu); b->stop = true; }
        b->cv_work.notify_all();
        for (auto &t : b->workers) t.join();
        delete b;
    } catch (...) {
    }
}

int ucfp_batcher_query(ucfp_batcher *b, const void *query, size_t k, uint64_t *ids_out, void *keys_out) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(b != nullptr, UCFP_E_INVALID, "null batcher");
    if (k == 0) return UCFP_OK;
    UCFP_REQUIRE(query && ids_out && keys_out, UCFP_E_INVALID, "NULL query or output buffer");
    Request r;
    r.query = query; r.k = k; r.ids_out = ids_out; r.keys_out = keys_out;
    r.t_in = std::chrono::steady_clock::now();
    {
        std::unique_lock<std::mutex> lk(b->mu);
        UCFP_REQUIRE(!b->stop, UCFP_E_STATE, "batcher is shutting down");
        b->queue.push_back(&r);
        if (b->queue.size() == 1 || b->queue.size() >= b->max_batch) b->cv_work.notify_one();
        b->cv_done.wait(lk, [&] { return r.done; });
    }
    if (r.rc != UCFP_OK) ucfp::set_error("%s", r.err);
    return r.rc;
    UCFP_API_END
}

int ucfp_batcher_stats(const ucfp_batcher *b, uint64_t *queries, uint64_t *batches, uint64_t *largest_batch) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(b != nullptr, UCFP_E_INVALID, "null batcher");
    if (queries) *queries = b->n_queries.load();
    if (batches) *batches = b->n_batches.load();
    if (largest_batch) *largest_batch = b->largest.load();
    return UCFP_OK;
    UCFP_API_END
}

}  // extern "C"
