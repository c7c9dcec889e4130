// corpus.cu -- mutation of an HBM-resident corpus: allocation, growth, delete and insert-or-replace by record id.
//
// Reference interface (paths relative to the reference tree): IndexBackend::upsert / ::delete, src/index/mod.rs:20-25
// ("Insert-or-replace by (tenant_id, record_id)", "Idempotent -- missing IDs are not an error"); the embedded backend
// implements them as redb writes (src/index/embedded/mod.rs:157-266).  Here the corpus is the HBM mirror of those tables:
//   delete   the rows of the given ids are found on the device (one pass over the id column, 8 B/row), the LAST rows of
//            the corpus move into the freed slots and the side arrays of the moved rows are rebuilt -- the corpus stays
//            dense, so no scan kernel ever sees a tombstone and scan results depend on record ids only, never on row
//            order (the total orders of all three scans break ties by record id);
//   upsert   rows whose id is resident are overwritten in place, the others are appended (the corpus grows by
//            reallocation + device-to-device copy when it is full).
// An implicit-id corpus (record id = id_base + row) becomes an explicit-id corpus the first time it is mutated this way.
#include <algorithm>
#include <unordered_map>

#include "api_util.cuh"

namespace ucfp {
namespace {

constexpr uint32_t kFilterBits = 1u << 15;   // 4 KiB of shared memory: Bloom-style prefilter of the target ids

__device__ __forceinline__ uint32_t id_hash(uint64_t id) { return (uint32_t)(mix64(id) >> 40) & (kFilterBits - 1); }

// out_rows[j] = row holding targets[j] (targets sorted ascending, unique), left untouched when absent.
__global__ void __launch_bounds__(256) find_rows_kernel(const uint64_t *__restrict__ ids, uint64_t n_rows, const uint64_t *__restrict__ targets,
                                                         uint32_t n_targets, unsigned long long *out_rows) {
    __shared__ uint32_t filter[kFilterBits / 32];
    for (uint32_t i = threadIdx.x; i < kFilterBits / 32; i += blockDim.x) filter[i] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_targets; i += blockDim.x) { const uint32_t h = id_hash(targets[i]); atomicOr(&filter[h >> 5], 1u << (h & 31)); }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const uint64_t id = ids[r];
        const uint32_t h = id_hash(id);
        if (!(filter[h >> 5] >> (h & 31) & 1u)) continue;
        uint32_t lo = 0, hi = n_targets;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (targets[mid] < id) lo = mid + 1; else hi = mid; }
        if (lo < n_targets && targets[lo] == id) out_rows[lo] = r;
    }
}

__global__ void fill_ids_kernel(uint64_t *ids, uint64_t id_base, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) ids[r] = id_base + r;
}

// One CTA per job: corpus row dst[j] <- row src[j] of `from` (words of 8 bytes); ids likewise when id_from is given,
// else ids[dst[j]] = new_ids[j].
__global__ void __launch_bounds__(128) copy_rows_kernel(uint32_t *rows, uint64_t *ids, const uint32_t *__restrict__ from, const uint64_t *id_from,
                                                         const uint64_t *new_ids, const uint64_t *__restrict__ dst, const uint64_t *__restrict__ src,
                                                         uint32_t words_per_row) {   // 32-bit words: every row size is a multiple of 4 (cosine rows are 4 x dim bytes)
    const uint64_t d = dst[blockIdx.x], s = src[blockIdx.x];
    for (uint32_t w = threadIdx.x; w < words_per_row; w += blockDim.x) rows[d * words_per_row + w] = from[s * words_per_row + w];
    if (threadIdx.x == 0 && ids) ids[d] = id_from ? id_from[s] : new_ids[blockIdx.x];
}

int to_host_ids(ucfp_lane *ln, const uint64_t *ids, uint64_t n, std::vector<uint64_t> &out) {
    out.resize(n);
    if (classify(ids) == Mem::Device) {
        UCFP_CUDA_TRY(cudaMemcpyAsync(out.data(), ids, 8 * n, cudaMemcpyDeviceToHost, ln->stream));
        UCFP_CUDA_TRY(cudaStreamSynchronize(ln->stream));
    } else {
        memcpy(out.data(), ids, 8 * n);
    }
    return UCFP_OK;
}

// record ids -> rows (UINT64_MAX when absent); `targets` sorted ascending and unique
int find_rows(ucfp_lane *ln, ucfp_corpus *c, const std::vector<uint64_t> &targets, std::vector<uint64_t> &rows) {
    const size_t n = targets.size();
    rows.assign(n, UINT64_MAX);
    if (n == 0 || c->size == 0) return UCFP_OK;
    if (c->id_mode != 1) {   // implicit ids: arithmetic
        for (size_t j = 0; j < n; ++j)
            if (targets[j] >= c->id_base && targets[j] - c->id_base < c->size) rows[j] = targets[j] - c->id_base;
        return UCFP_OK;
    }
    UCFP_TRY(ln->misc.reserve(16 * n));
    uint64_t *t_dev = ln->misc.as<uint64_t>();
    unsigned long long *r_dev = reinterpret_cast<unsigned long long *>(t_dev + n);
    UCFP_CUDA_TRY(cudaMemcpyAsync(t_dev, targets.data(), 8 * n, cudaMemcpyHostToDevice, ln->stream));
    UCFP_CUDA_TRY(cudaMemsetAsync(r_dev, 0xFF, 8 * n, ln->stream));
    constexpr size_t kChunk = 1u << 14;   // targets per launch: keeps the prefilter selective
    for (size_t lo = 0; lo < n; lo += kChunk) {
        const uint32_t m = (uint32_t)std::min(kChunk, n - lo);
        uint64_t blocks = (c->size + 255) / 256;
        const uint64_t maxb = (uint64_t)ln->sm_count * 8;
        if (blocks > maxb) blocks = maxb;
        find_rows_kernel<<<(unsigned)blocks, 256, 0, ln->stream>>>(c->ids, c->size, t_dev + lo, m, r_dev + lo);
        count_launch(ln);
    }
    UCFP_TRY(check_launch("find_rows"));
    UCFP_CUDA_TRY(cudaMemcpyAsync(rows.data(), r_dev, 8 * n, cudaMemcpyDeviceToHost, ln->stream));
    UCFP_CUDA_TRY(cudaStreamSynchronize(ln->stream));
    return UCFP_OK;
}

int materialise_ids(ucfp_lane *ln, ucfp_corpus *c) {
    if (c->id_mode == 1) return UCFP_OK;
    if (!c->ids) UCFP_CUDA_TRY(cudaMalloc((void **)&c->ids, 8 * (c->capacity + 16)));
    if (c->size) {
        uint64_t blocks = std::min<uint64_t>((c->size + 255) / 256, (uint64_t)ln->sm_count * 8);
        fill_ids_kernel<<<(unsigned)blocks, 256, 0, ln->stream>>>(c->ids, c->id_base, c->size);
        count_launch(ln);
        UCFP_TRY(check_launch("fill_ids"));
    }
    c->id_mode = 1;
    return UCFP_OK;
}

// side arrays of rows rewritten in place: one by one for a handful, everything otherwise
int rebuild_rows(ucfp_lane *ln, ucfp_corpus *c, const std::vector<uint64_t> &touched) {
    if (touched.empty() || c->size == 0) return UCFP_OK;
    if (touched.size() > 64) return after_append(ln, c, 0, c->size);
    for (uint64_t r : touched)
        if (r < c->size) UCFP_TRY(after_append(ln, c, r, 1));
    return UCFP_OK;
}

int launch_copy_rows(ucfp_lane *ln, ucfp_corpus *c, const void *from, const uint64_t *id_from, const std::vector<uint64_t> &new_ids,
                     const std::vector<uint64_t> &dst, const std::vector<uint64_t> &src) {
    const size_t n = dst.size();
    if (n == 0) return UCFP_OK;
    const size_t lists = new_ids.empty() ? 2 : 3;
    UCFP_TRY(ln->flags.reserve(8 * n * lists));
    uint64_t *d_dev = ln->flags.as<uint64_t>(), *s_dev = d_dev + n, *i_dev = new_ids.empty() ? nullptr : s_dev + n;
    UCFP_CUDA_TRY(cudaMemcpyAsync(d_dev, dst.data(), 8 * n, cudaMemcpyHostToDevice, ln->stream));
    UCFP_CUDA_TRY(cudaMemcpyAsync(s_dev, src.data(), 8 * n, cudaMemcpyHostToDevice, ln->stream));
    if (i_dev) UCFP_CUDA_TRY(cudaMemcpyAsync(i_dev, new_ids.data(), 8 * n, cudaMemcpyHostToDevice, ln->stream));
    for (size_t lo = 0; lo < n; lo += 65535) {   // grid.x limit is far away, but keep launches modest
        const unsigned m = (unsigned)std::min<size_t>(65535, n - lo);
        copy_rows_kernel<<<m, 128, 0, ln->stream>>>(static_cast<uint32_t *>(c->rows), c->ids, static_cast<const uint32_t *>(from), id_from,
                                                    i_dev ? i_dev + lo : nullptr, d_dev + lo, s_dev + lo, (uint32_t)(row_bytes(c) / 4));
        count_launch(ln);
    }
    UCFP_TRY(check_launch("copy_rows"));
    // the host vectors die with the caller's frame: the uploads above must have been consumed
    UCFP_CUDA_TRY(cudaStreamSynchronize(ln->stream));
    return UCFP_OK;
}

}  // namespace

int corpus_device_init(ucfp_ctx *) { return UCFP_OK; }

int corpus_alloc_arrays(ucfp_lane *ln, ucfp_corpus *c, uint64_t capacity) {
    const int kind = c->kind;
    const size_t rb = row_bytes(c);
    // +16 rows of slack so that vector loads of the last partial tile never leave the allocation
    cudaError_t e = cudaMalloc(&c->rows, rb * (capacity + 16));
    if (e == cudaSuccess && kind == UCFP_KIND_MINHASH128) e = cudaMalloc((void **)&c->mh_sketch, 2 * 128 * ((capacity + 31) / 32 * 32 + 512));  // two planes (byte 0, byte 1 of every slot); whole 256-row tiles stay readable
    if (e == cudaSuccess && kind == UCFP_KIND_COSINE) {
        c->dim_pad = (c->dim + 63) / 64 * 64;
        e = cudaMalloc(&c->cos_bf16, 2 * (size_t)c->dim_pad * (capacity + 256));
        if (e == cudaSuccess) e = cudaMalloc((void **)&c->cos_inv_norm, 4 * (capacity + 256));
    }
    if (e == cudaSuccess && kind == UCFP_KIND_HAMMING64) {
        // Operand rows of the tensor-core scan, 32 B per code on top of the 8 B code.  Optional: without them (allocation
        // refused) the scan expands the codes on the fly in its producer warps, slower for 64-512-query batches.
        const size_t ops_bytes = 64 * ((capacity + 1) / 2 + 512);   // whole 256-row stages stay readable
        if (cudaMalloc((void **)&c->ham_ops, ops_bytes) != cudaSuccess) { cudaGetLastError(); c->ham_ops = nullptr; }
        else if (cudaMemsetAsync(c->ham_ops, 0, ops_bytes, ln->stream) != cudaSuccess) { cudaGetLastError(); cudaFree(c->ham_ops); c->ham_ops = nullptr; }
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("corpus allocation of %llu rows failed: %s", (unsigned long long)capacity, cudaGetErrorString(e));
        if (c->rows) cudaFree(c->rows);
        if (c->mh_sketch) cudaFree(c->mh_sketch);
        if (c->cos_bf16) cudaFree(c->cos_bf16);
        if (c->cos_inv_norm) cudaFree(c->cos_inv_norm);
        if (c->ham_ops) cudaFree(c->ham_ops);
        c->rows = nullptr; c->mh_sketch = nullptr; c->cos_bf16 = nullptr; c->cos_inv_norm = nullptr; c->ham_ops = nullptr;
        return UCFP_E_OOM;
    }
    if (kind == UCFP_KIND_MULTIHASH) {   // side corpus of the PHash global hashes, scanned by the coarse pass of a re-rank
        c->coarse = new (std::nothrow) ucfp_corpus();
        int rc = c->coarse ? UCFP_OK : UCFP_E_OOM;
        if (rc == UCFP_OK) { c->coarse->ctx = c->ctx; c->coarse->kind = UCFP_KIND_HAMMING64; rc = corpus_alloc_arrays(ln, c->coarse, capacity); }
        if (rc != UCFP_OK) { delete c->coarse; c->coarse = nullptr; cudaFree(c->rows); c->rows = nullptr; return rc; }
        c->coarse->id_mode = 2;
    }
    c->capacity = capacity;
    return UCFP_OK;
}

// Reallocates every array for `capacity` rows, copies rows and ids device to device, re-derives the side arrays.
int corpus_grow(ucfp_lane *ln, ucfp_corpus *c, uint64_t capacity) {
    if (capacity <= c->capacity) return UCFP_OK;
    ucfp_corpus fresh;
    fresh.kind = c->kind; fresh.dim = c->dim; fresh.ctx = c->ctx;
    UCFP_TRY(corpus_alloc_arrays(ln, &fresh, capacity));
    const size_t rb = row_bytes(c);
    cudaError_t e = cudaSuccess;
    if (c->size) e = cudaMemcpyAsync(fresh.rows, c->rows, rb * c->size, cudaMemcpyDeviceToDevice, ln->stream);
    if (e == cudaSuccess && c->ids) {
        e = cudaMalloc((void **)&fresh.ids, 8 * (capacity + 16));
        if (e == cudaSuccess && c->size) e = cudaMemcpyAsync(fresh.ids, c->ids, 8 * c->size, cudaMemcpyDeviceToDevice, ln->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ln->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("growing the corpus to %llu rows failed: %s", (unsigned long long)capacity, cudaGetErrorString(e));
        void *arrs[] = {fresh.rows, fresh.ids, fresh.ham_ops, fresh.mh_sketch, fresh.cos_bf16, fresh.cos_inv_norm,
                        fresh.coarse ? fresh.coarse->rows : nullptr, fresh.coarse ? (void *)fresh.coarse->ham_ops : nullptr};
        for (void *a : arrs) if (a) cudaFree(a);
        delete fresh.coarse;
        return e == cudaErrorMemoryAllocation ? UCFP_E_OOM : UCFP_E_CUDA;
    }
    void *old[] = {c->rows, c->ids, c->ham_ops, c->mh_sketch, c->cos_bf16, c->cos_inv_norm,
                   c->coarse ? c->coarse->rows : nullptr, c->coarse ? (void *)c->coarse->ham_ops : nullptr};
    for (void *a : old) if (a) cudaFree(a);
    delete c->coarse;
    c->coarse = fresh.coarse;
    c->rows = fresh.rows; c->ids = fresh.ids; c->ham_ops = fresh.ham_ops; c->mh_sketch = fresh.mh_sketch;
    c->cos_bf16 = fresh.cos_bf16; c->cos_inv_norm = fresh.cos_inv_norm; c->dim_pad = fresh.dim_pad; c->capacity = capacity;
    return after_append(ln, c, 0, c->size);
}

int corpus_delete_ids(ucfp_lane *ln, ucfp_corpus *c, const uint64_t *ids, uint64_t n, uint64_t *n_removed) {
    std::vector<uint64_t> targets;
    UCFP_TRY(to_host_ids(ln, ids, n, targets));
    std::sort(targets.begin(), targets.end());
    targets.erase(std::unique(targets.begin(), targets.end()), targets.end());
    std::vector<uint64_t> rows;
    UCFP_TRY(find_rows(ln, c, targets, rows));
    std::vector<uint64_t> dead;
    for (uint64_t r : rows) if (r != UINT64_MAX) dead.push_back(r);
    if (dead.empty()) return UCFP_OK;   // idempotent: unknown ids are not an error (src/index/mod.rs:23-25)
    std::sort(dead.begin(), dead.end());
    UCFP_TRY(materialise_ids(ln, c));
    const uint64_t new_size = c->size - dead.size();
    // holes below new_size are filled by the live rows at or above it
    std::vector<uint64_t> dst, src;
    size_t di = std::lower_bound(dead.begin(), dead.end(), new_size) - dead.begin();   // dead rows in the tail start here
    uint64_t tail = new_size;
    for (size_t h = 0; h < dead.size() && dead[h] < new_size; ++h) {
        while (di < dead.size() && dead[di] == tail) { ++di; ++tail; }
        dst.push_back(dead[h]);
        src.push_back(tail++);
    }
    UCFP_TRY(launch_copy_rows(ln, c, c->rows, c->ids, {}, dst, src));
    c->size = new_size;
    if (new_size == 0) c->id_mode = 1;
    UCFP_TRY(rebuild_rows(ln, c, dst));
    // the pair row of the new last code may still carry a partner that is now beyond the end: rebuild it (Hamming)
    if (c->kind == UCFP_KIND_HAMMING64 && new_size) UCFP_TRY(after_append(ln, c, new_size - 1, 1));
    if (n_removed) *n_removed = dead.size();
    return UCFP_OK;
}

int corpus_upsert_rows(ucfp_lane *ln, ucfp_corpus *c, const uint64_t *ids, const void *rows, uint64_t n, uint64_t *n_replaced) {
    std::vector<uint64_t> hid;
    UCFP_TRY(to_host_ids(ln, ids, n, hid));
    for (uint64_t id : hid) UCFP_REQUIRE(id != UCFP_ID_NONE, UCFP_E_INVALID, "UCFP_ID_NONE is not a valid record id");
    // within one batch the last occurrence of an id wins, as consecutive redb inserts would
    std::unordered_map<uint64_t, uint64_t> last;
    last.reserve(n * 2);
    for (uint64_t j = 0; j < n; ++j) last[hid[j]] = j;
    std::vector<uint64_t> targets;
    targets.reserve(last.size());
    for (auto &kv : last) targets.push_back(kv.first);
    std::sort(targets.begin(), targets.end());
    UCFP_TRY(materialise_ids(ln, c));
    std::vector<uint64_t> found;
    UCFP_TRY(find_rows(ln, c, targets, found));
    std::vector<uint64_t> dst, src, new_ids, touched;
    uint64_t n_new = 0;
    for (size_t t = 0; t < targets.size(); ++t) {
        src.push_back(last[targets[t]]);
        new_ids.push_back(targets[t]);
        if (found[t] != UINT64_MAX) { dst.push_back(found[t]); touched.push_back(found[t]); }
        else dst.push_back(c->size + n_new++);
    }
    if (c->size + n_new > c->capacity) {
        uint64_t want = c->capacity * 2;
        if (want < c->size + n_new) want = c->size + n_new;
        UCFP_TRY(corpus_grow(ln, c, want));
    }
    // stage the batch on the device (host rows), then one kernel scatters rows and ids into place
    const size_t rb = row_bytes(c);
    const void *from = rows;
    if (classify(rows) != Mem::Device) {
        UCFP_TRY(ln->cand.reserve(rb * n));
        UCFP_CUDA_TRY(cudaMemcpyAsync(ln->cand.ptr, rows, rb * n, cudaMemcpyHostToDevice, ln->stream));
        from = ln->cand.ptr;
    }
    UCFP_TRY(launch_copy_rows(ln, c, from, nullptr, new_ids, dst, src));
    const uint64_t first_new = c->size;
    c->size += n_new;
    c->id_mode = 1;
    if (n_new) UCFP_TRY(after_append(ln, c, first_new, n_new));
    UCFP_TRY(rebuild_rows(ln, c, touched));
    if (n_replaced) *n_replaced = touched.size();
    return UCFP_OK;
}

}  // namespace ucfp
