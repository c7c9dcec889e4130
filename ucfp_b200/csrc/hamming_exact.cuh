// hamming_exact.cuh -- unconditional exact selection for queries whose candidate list
// overflowed in the threshold scan (included by hamming.cu inside namespace ucfp).
//
// One cooperative launch per query pass; when no query is flagged every CTA leaves before
// the first grid barrier, so the common case costs one empty launch and no host sync.
// For each flagged query, with bounded memory and for ANY input:
//   1. histogram of dist over the whole corpus            -> d* = distance of the k-th result,
//                                                             need = how many rows at d* belong
//   2. 8 x 8-bit MSB-first radix select over the ids of rows with dist == d*
//                                                          -> id* = need-th smallest such id
//   3. collect rows with dist < d* or (dist == d* and id <= id*)   (exactly k rows)
//   4. CTA 0 sorts the k rows by (dist, id) and writes the result slots
// That is 10 HBM passes per flagged query; it only runs for pathological corpora
// (more than `cap` rows tying inside the top-k window in descending-id order).


namespace {

namespace cg = cooperative_groups;

struct ExactScratch {
    unsigned long long hist[66];     // distances 0..64
    unsigned long long digit[256];   // radix pass histogram
    unsigned int out_count;
    unsigned int pad;
};

__device__ __forceinline__ uint32_t ham64(uint64_t a, uint32_t lo, uint32_t hi) {
    return __popc((uint32_t)a ^ lo) + __popc((uint32_t)(a >> 32) ^ hi);
}

__global__ void __launch_bounds__(256)
hamming_exact_kernel(const uint64_t *__restrict__ codes, const uint64_t *__restrict__ ids, uint64_t id_base, uint64_t N,
                     const QSlot *__restrict__ slots, const uint32_t *__restrict__ flags, uint32_t nq, uint32_t k,
                     ExactScratch *scr, uint64_t *out_id, uint32_t *out_d, uint64_t *ids_out, uint32_t *dist_out) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned int s_any;
    __shared__ unsigned long long s_hist[256];
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nq; i += blockDim.x) if (flags[i]) s_any = 1;
    __syncthreads();
    if (!s_any) return;  // uniform over the grid: every CTA reads the same flags

    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t gsize = (uint64_t)gridDim.x * blockDim.x;

    for (uint32_t q = 0; q < nq; ++q) {
        if (!flags[q]) continue;
        const QSlot s = slots[q];
        // ---- 1. distance histogram
        if (gtid < 66) scr->hist[gtid] = 0;
        if (gtid == 0) scr->out_count = 0;
        if (threadIdx.x < 66) s_hist[threadIdx.x] = 0;
        grid.sync();
        for (uint64_t r = gtid; r < N; r += gsize) atomicAdd(&s_hist[ham64(codes[r], s.lo, s.hi)], 1ULL);
        __syncthreads();
        if (threadIdx.x < 66 && s_hist[threadIdx.x]) atomicAdd(&scr->hist[threadIdx.x], s_hist[threadIdx.x]);
        grid.sync();
        uint32_t dstar = 65; uint64_t need = 0;
        {
            uint64_t cum = 0;
            for (uint32_t d = 0; d <= 64; ++d) {
                uint64_t h = scr->hist[d];
                if (cum + h >= k) { dstar = d; need = k - cum; break; }
                cum += h;
            }
        }
        // fewer than k rows in total: take everything
        uint64_t idstar = UINT64_MAX;
        if (dstar <= 64) {
            // ---- 2. radix select of the need-th smallest id among rows at distance d*
            uint64_t prefix = 0; uint64_t want = need;  // 1-based rank inside the current prefix bucket
            for (int shift = 56; shift >= 0; shift -= 8) {
                if (gtid < 256) scr->digit[gtid] = 0;
                s_hist[threadIdx.x] = 0;
                grid.sync();
                const uint64_t hi_mask = shift == 56 ? 0 : ~0ULL << (shift + 8);
                for (uint64_t r = gtid; r < N; r += gsize) {
                    if (ham64(codes[r], s.lo, s.hi) != dstar) continue;
                    uint64_t id = ids ? ids[r] : id_base + r;
                    if ((id & hi_mask) != prefix) continue;
                    atomicAdd(&s_hist[(id >> shift) & 255], 1ULL);
                }
                __syncthreads();
                if (s_hist[threadIdx.x]) atomicAdd(&scr->digit[threadIdx.x], s_hist[threadIdx.x]);
                grid.sync();
                uint64_t cum = 0; uint32_t dig = 255;
                for (uint32_t b = 0; b < 256; ++b) {
                    uint64_t h = scr->digit[b];
                    if (cum + h >= want) { dig = b; break; }
                    cum += h;
                }
                want -= cum;
                prefix |= (uint64_t)dig << shift;
                grid.sync();  // everyone has read digit[] before it is cleared again
            }
            idstar = prefix;
        }
        // ---- 3. collect
        for (uint64_t r = gtid; r < N; r += gsize) {
            uint32_t d = ham64(codes[r], s.lo, s.hi);
            if (d > dstar) continue;
            uint64_t id = ids ? ids[r] : id_base + r;
            if (d == dstar && id > idstar) continue;
            unsigned int pos = atomicAdd(&scr->out_count, 1u);
            if (pos < k) { out_id[pos] = id; out_d[pos] = d; }
        }
        grid.sync();
        // ---- 4. CTA 0 orders the winners (k <= 2048: rank by counting, O(k^2) on a tiny set)
        if (blockIdx.x == 0) {
            uint32_t m = min(scr->out_count, k);
            for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
                if (i >= m) { ids_out[(size_t)q * k + i] = UINT64_MAX; dist_out[(size_t)q * k + i] = UINT32_MAX; }
            }
            for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                uint64_t id = out_id[i]; uint32_t d = out_d[i];
                uint32_t rank = 0;
                for (uint32_t j = 0; j < m; ++j) {
                    uint64_t idj = out_id[j]; uint32_t dj = out_d[j];
                    rank += (dj < d || (dj == d && (idj < id || (idj == id && j < i))));
                }
                ids_out[(size_t)q * k + rank] = id; dist_out[(size_t)q * k + rank] = d;
            }
        }
        grid.sync();
    }
}

}  // namespace

static int hamming_exact_fallback(ucfp_corpus *c, const QSlot *slots, const uint32_t *flags, uint32_t nq, uint32_t k,
                                  uint64_t *ids_out, uint32_t *dist_out) {
    ucfp_ctx *ctx = c->ctx;
    size_t scratch = sizeof(ExactScratch) + (sizeof(uint64_t) + sizeof(uint32_t)) * (size_t)k + 64;
    UCFP_TRY(ctx->misc.reserve(scratch));
    ExactScratch *scr = ctx->misc.as<ExactScratch>();
    uint64_t *out_id = reinterpret_cast<uint64_t *>(scr + 1);
    uint32_t *out_d = reinterpret_cast<uint32_t *>(out_id + k);
    int occ = 0;
    UCFP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hamming_exact_kernel, 256, 0));
    if (occ < 1) occ = 1;
    if (occ > 4) occ = 4;
    const uint64_t *codes = static_cast<const uint64_t *>(c->rows);
    const uint64_t *ids = c->id_mode == 1 ? c->ids : nullptr;
    uint64_t id_base = c->id_base, N = c->size;
    void *args[] = {&codes, &ids, &id_base, &N, &slots, &flags, &nq, &k, &scr, &out_id, &out_d, &ids_out, &dist_out};
    UCFP_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)hamming_exact_kernel, dim3(ctx->sm_count * occ), dim3(256), args, 0,
                                              ctx->stream));
    count_launch(ctx);
    return UCFP_OK;
}
