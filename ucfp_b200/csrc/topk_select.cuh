// topk_select.cuh -- the selection machinery shared by the integer-keyed scans (Hamming: key = distance,
// Jaccard: key = 128 - matches; smaller key is better, ties by record id ascending).
//
//   compact_kernel       one CTA per query: bitonic sort of the candidate list by (key, id), keep k, publish
//                        the new admission bound (thr = key_k, kth_id = id_k); the last call writes results.
//   re-scan rounds       a query whose list overflowed between two compactions (floods of equal keys arriving in descending id
//                        order) is scanned again, alone with the other flagged queries, over the whole corpus under the bound
//                        its truncated lists produced: that bound is the k-th best of real rows, hence an upper bound of the
//                        true k-th best, and far fewer rows pass it (each round shrinks the flood by ~0.75 cap / k).
//   exact_select_kernel  cooperative multi-pass exact selection (key histogram + 8-bit radix select on ids)
//                        for queries still flagged after the re-scan rounds; correct for ANY input with bounded memory.
//
// Included inside namespace ucfp { namespace { ... } } by each scan's .cu file.
#pragma once

constexpr uint64_t kRowMask = (1ULL << 40) - 1;  // candidate = key << 40 | row

struct SelectState {      // per query pass, all device pointers
    uint64_t *cand;       // [nq][cap] candidate entries
    uint32_t *count;      // [nq] entries appended (may exceed cap: overflow)
    uint32_t *thr;        // first query's admission bound; query q at thr[q * thr_stride]
    uint32_t thr_stride;
    uint64_t *kth_id;     // [nq]
    uint32_t *flags;      // [nq] 1 = list overflowed: the query is re-scanned (kRescanRounds), then left to exact_select
    uint32_t cap;
    uint32_t *big;        // [nq] list too long for the small compaction launch
    unsigned long long *max_fill = nullptr;   // diagnostics (may be null): longest list any compaction of this scan has seen
    int emit_rows = 0;    // the final pass reports ROW indices instead of record ids (ties are still broken by record id): the coarse
                          // pass of the multi-hash re-rank needs the rows, and its candidate set must not depend on the row order
    int small_keys = 0;   // keys are small integers (Hamming distance, 128 - MinHash matches): compaction first drops, by a 256-bin
                          // histogram of the keys, every entry beyond the k-th key, and sorts the few that are left
};

// Every CTA of a re-scan launch derives the same list of flagged queries (ascending): warp 0 calls this, emit(position, query)
// stores whatever the scan needs per listed query; returns the length of the list.
template <typename Emit>
__device__ __forceinline__ uint32_t list_flagged_queries(const uint32_t *__restrict__ flags, uint32_t nq, Emit emit) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    for (uint32_t q0 = 0; q0 < nq; q0 += 32) {
        const uint32_t q = q0 + lane;
        const bool f = q < nq && flags[q] != 0;
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (f) emit(base + (uint32_t)__popc(b & ((1u << lane) - 1u)), q);
        base += (uint32_t)__popc(b);
    }
    return base;
}

constexpr uint32_t kSmallList = 1024;   // entries the small compaction launch sorts (16 KiB of shared memory)

// Appends one entry to query q's candidate list.  A full list does not simply drop the newcomer: positions [cap / 4, cap) are a
// RESERVOIR (Algorithm R with a hash of the arrival number as its random source), so that what a flooded list holds at the next
// compaction is a uniform sample of everything that was admitted, whatever the order of arrival.  Neither "first come" nor "last
// writer wins" is good enough: under a flood of equal keys whose ids descend with the row the first arrivals are the worst rows, and
// the last ones are the stragglers of the launch's first wave, not its best rows -- a "last writer wins" list produced bounds that still
// admitted 100 K of 600 K flood rows.  The k-th best of a uniform sample of M admitted rows has rank ~ k M / (0.75 cap) among them; that
// is what makes the re-scan rounds converge.  The first quarter of the list (the k entries kept by the previous compaction, k <= cap / 4,
// and the earliest arrivals) is never overwritten.  cap is a power of two; an entry is one 64-bit word, so racing writers leave one of
// their entries, never a torn one.
__device__ __forceinline__ void cand_append(uint64_t *cand, uint32_t *count, uint32_t cap, uint32_t q, uint64_t entry) {
    uint32_t pos = atomicAdd(&count[q], 1u);
    if (pos >= cap) {
        const uint32_t lo = cap >> 2;
        uint32_t h = pos * 0x9E3779B1u;
        h ^= h >> 15; h *= 0x85EBCA6Bu; h ^= h >> 13;
        const uint32_t j = __umulhi(h, pos - lo + 1u);   // uniform over the arrivals that competed for the reservoir so far
        if (j >= cap - lo) return;
        pos = lo + j;
    }
    cand[(size_t)q * cap + pos] = entry;
}

__device__ __forceinline__ bool cand_before(uint64_t dra, uint64_t ida, uint64_t drb, uint64_t idb) {
    uint32_t da = (uint32_t)(dra >> 40), db = (uint32_t)(drb >> 40);
    return da < db || (da == db && ida < idb);
}

// key_flip: 0 -> the reported value is the key itself; otherwise reported = key_flip - key (Jaccard matches).
// Two launches per step (compact_lists): the first with shared memory for kSmallList entries (many CTAs per SM, one wave)
// handles the usual short lists and leaves longer ones, marked in S.big, to the second launch (16 B x cap of shared memory),
// whose other CTAs return at once.  second_launch == 2 is the compaction of a re-scan round: it takes exactly the flagged queries,
// whatever the length of their lists, and clears the flag of every query whose re-scan fitted.
constexpr int kRescanRounds = 2;   // each round divides a flood by ~0.75 cap / k.  (Dropping a flagged query from the tensor scan's hot test for the rest of the
                                   // main scan was measured too: the first round then meets the whole flood, a third round is needed, and the batch is no faster:
                                   // 8.65 vs 9.56 ms with 8 of 1 024 queries in a 1 M-row flood, +40 us on every batch that has none.)
static inline int rescan_rounds() {   // developer switch UCFP_RESCAN_ROUNDS (read once): 0 sends every overflow straight to exact_select
    static const int rounds = getenv("UCFP_RESCAN_ROUNDS") ? atoi(getenv("UCFP_RESCAN_ROUNDS")) : kRescanRounds;
    return rounds < 0 ? 0 : (rounds > 8 ? 8 : rounds);
}
__global__ void compact_kernel(SelectState S, uint32_t k, const uint64_t *__restrict__ ids, uint64_t id_base, int final_pass,
                               uint32_t key_flip, uint64_t *ids_out, uint32_t *keys_out, uint32_t smem_entries, int second_launch) {
    extern __shared__ uint64_t sm[];
    const uint32_t q = blockIdx.x;
    const uint32_t n_raw = S.count[q];
    const uint32_t n = min(n_raw, S.cap);
    if (!second_launch && threadIdx.x == 0 && S.max_fill && n_raw > 64) atomicMax(S.max_fill, (unsigned long long)n_raw);   // short lists: not worth an atomic
    if (second_launch == 2) {
        if (!S.flags[q]) return;
    } else if (!second_launch) {
        const bool big = n > smem_entries;
        if (threadIdx.x == 0) S.big[q] = big ? 1u : 0u;
        if (big) return;
    } else if (!S.big[q]) return;
    uint64_t *list = S.cand + (size_t)q * S.cap;
    // Small integer keys: the entries that can be among the k best are those whose key is <= the k-th smallest KEY, which a
    // histogram finds without sorting.  Of a seed list of 1024 random codes ~20 survive (k = 10), so the bitonic network below
    // runs on 32 entries instead of 1024 (the seed compaction of a 1024-query batch was 91 us of a 5 ms shard scan).
    __shared__ uint32_t s_hist[256], s_cut, s_keep, s_fill;
    uint32_t n_keep = n, cut = 0xFFFFFFFFu;
    if (S.small_keys && n > 32 && n > 2 * k) {
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        if (threadIdx.x == 0) s_fill = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&s_hist[min((uint32_t)(list[i] >> 40), 255u)], 1u);
        __syncthreads();
        if (threadIdx.x < 32) {   // bin of the k-th smallest key and the number of entries up to and including that bin
            const uint32_t lane = threadIdx.x;
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { loc[j] = s_hist[lane * 8 + j]; sum += loc[j]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
            uint32_t cum = incl - sum;
            if (cum < k && k <= incl) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { cum += loc[j]; if (cum >= k) { s_cut = lane * 8 + j; s_keep = cum; break; } }
            }
            if (lane == 31 && incl < k) { s_cut = 255; s_keep = incl; }   // fewer than k entries: keep all (cannot happen with n > 2k)
        }
        __syncthreads();
        cut = s_cut; n_keep = s_keep;
    }
    uint32_t P = 1;
    while (P < n_keep) P <<= 1;
    uint64_t *s_id = sm, *s_dr = sm + P;
    if (cut != 0xFFFFFFFFu) {
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) { s_id[i] = UINT64_MAX; s_dr[i] = UINT64_MAX; }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t e = list[i];
            if (min((uint32_t)(e >> 40), 255u) <= cut) {   // slot order is arbitrary; the sort's order is total (ids are unique)
                const uint32_t slot = atomicAdd(&s_fill, 1u);
                const uint64_t r = e & kRowMask;
                s_dr[slot] = e; s_id[slot] = ids ? ids[r] : id_base + r;
            }
        }
    } else {
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
            uint64_t e = UINT64_MAX, id = UINT64_MAX;
            if (i < n) { e = list[i]; uint64_t r = e & kRowMask; id = ids ? ids[r] : id_base + r; }
            s_id[i] = id; s_dr[i] = e;
        }
    }
    __syncthreads();
    for (uint32_t size = 2; size <= P; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                uint32_t i = 2 * t - (t & (stride - 1));  // lower index of the pair
                uint32_t j = i + stride;
                bool up = ((i & size) == 0);
                uint64_t di = s_dr[i], ii = s_id[i], dj = s_dr[j], ij = s_id[j];
                bool swap = up ? cand_before(dj, ij, di, ii) : cand_before(di, ii, dj, ij);
                if (swap) { s_dr[i] = dj; s_id[i] = ij; s_dr[j] = di; s_id[j] = ii; }
            }
            __syncthreads();
        }
    }
    const uint32_t m = min(n_keep, k);
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) list[i] = s_dr[i];
    if (final_pass) {
        for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
            bool ok = i < m;
            uint32_t key = ok ? (uint32_t)(s_dr[i] >> 40) : 0;
            ids_out[(size_t)q * k + i] = ok ? (S.emit_rows ? (s_dr[i] & kRowMask) : s_id[i]) : UINT64_MAX;
            keys_out[(size_t)q * k + i] = ok ? (key_flip ? key_flip - key : key) : UINT32_MAX;
        }
    }
    if (threadIdx.x == 0) {
        S.count[q] = m;
        if (second_launch == 2) S.flags[q] = n_raw > S.cap ? 1u : 0u;   // a re-scan that fitted is the query's exact answer
#ifdef UCFP_DEBUG_RESCAN
        if (second_launch == 2 && S.max_fill) { atomicMax(S.max_fill + 2, (unsigned long long)n_raw); atomicAdd(S.max_fill + 3, 1ULL); }
#endif
        else if (n_raw > S.cap) S.flags[q] = 1;
        // Publish the k-th entry as the admission bound -- unless the bound in force is already tighter: in a multi-GPU group
        // scan the ranks exchange their bounds between chunks (bounds_min_kernel below), so the bound a shard filters at may
        // come from another shard and be better than anything this list holds.
        if (n >= k) {
            const uint32_t new_thr = (uint32_t)(s_dr[k - 1] >> 40), cur_thr = S.thr[(size_t)q * S.thr_stride];
            const uint64_t new_kid = s_id[k - 1];
            if (new_thr < cur_thr || (new_thr == cur_thr && new_kid < S.kth_id[q])) { S.thr[(size_t)q * S.thr_stride] = new_thr; S.kth_id[q] = new_kid; }
        }
        // The scan is over and this query's lists lost entries on the way: what was written above is provisional.  Prepare its
        // re-scan: an empty list, and the bound made inclusive -- (thr, kth_id + 1) admits the bound's own row, so the rows of
        // the truncated lists need not be kept and nothing can enter the new list twice.
        if (final_pass && S.flags[q]) {
            S.count[q] = 0;
            if (S.kth_id[q] != UINT64_MAX) S.kth_id[q] += 1;
        }
    }
}

static inline void compact_rescanned(const SelectState &S, uint32_t nq, uint32_t k, const uint64_t *ids, uint64_t id_base, uint32_t key_flip,
                                     uint64_t *ids_out, uint32_t *keys_out, cudaStream_t st) {
    compact_kernel<<<nq, 512, 16 * (size_t)S.cap, st>>>(S, k, ids, id_base, 1, key_flip, ids_out, keys_out, S.cap, 2);
}

static inline void compact_lists(const SelectState &S, uint32_t nq, uint32_t k, const uint64_t *ids, uint64_t id_base, bool final_pass,
                                 uint32_t key_flip, uint64_t *ids_out, uint32_t *keys_out, cudaStream_t st) {
    compact_kernel<<<nq, 256, 16 * (size_t)kSmallList, st>>>(S, k, ids, id_base, final_pass ? 1 : 0, key_flip, ids_out, keys_out, kSmallList, 0);
    compact_kernel<<<nq, 512, 16 * (size_t)S.cap, st>>>(S, k, ids, id_base, final_pass ? 1 : 0, key_flip, ids_out, keys_out, S.cap, 1);
}

// ---- bound exchange between the shards of a group scan (group.cu) -------------------------------------------------
// Every rank holds, per query, the admission bound (thr, kth_id) = its current k-th best (key, record id): an upper bound
// of the GLOBAL k-th best.  The lexicographic minimum over the ranks is the tightest bound anyone knows; filtering every
// shard at it removes the cold-path work a small shard otherwise spends while its own bound is still loose.
struct __align__(16) BoundRec { uint64_t kth_id; uint32_t thr; uint32_t pad; };

__global__ void bounds_pack_kernel(const uint32_t *__restrict__ thr, uint32_t thr_stride, const uint64_t *__restrict__ kth_id, uint32_t nq, BoundRec *out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) out[q] = BoundRec{kth_id[q], thr[(size_t)q * thr_stride], 0u};
}
__global__ void bounds_min_kernel(const BoundRec *__restrict__ all, uint32_t world, uint32_t nq, uint32_t *thr, uint32_t thr_stride, uint64_t *kth_id) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint32_t t = thr[(size_t)q * thr_stride];
    uint64_t id = kth_id[q];
    for (uint32_t r = 0; r < world; ++r) {
        const BoundRec b = all[(size_t)r * nq + q];
        if (b.thr < t || (b.thr == t && b.kth_id < id)) { t = b.thr; id = b.kth_id; }
    }
    thr[(size_t)q * thr_stride] = t;
    kth_id[q] = id;
}
// One exchange on the lane's stream: pack, all-gather through the hook the group installed, fold the minimum back in.
static int exchange_bounds(ucfp_lane *ctx, const SelectState &S, uint32_t nq) {
    ucfp_exchange *x = ctx->xch;
    if (x->done >= x->max_real) { x->done++; return UCFP_OK; }   // the same on every rank (group.cu): collectives stay matched
    UCFP_TRY(x->send.reserve(sizeof(BoundRec) * nq));
    UCFP_TRY(x->recv.reserve(sizeof(BoundRec) * (size_t)nq * x->world));
    bounds_pack_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(S.thr, S.thr_stride, S.kth_id, nq, x->send.as<BoundRec>());
    UCFP_TRY(x->allgather(x->comm, x->send.ptr, x->recv.ptr, sizeof(BoundRec) * nq, ctx->stream));
    bounds_min_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(x->recv.as<BoundRec>(), (uint32_t)x->world, nq, S.thr, S.thr_stride, S.kth_id);
    count_launch(ctx, 2);
    x->done++;
    return UCFP_OK;
}
// Every rank performs exactly kBoundExchanges exchanges per query pass, whatever its shard size: after the chunks that end
// at or beyond these row counts, and any that are left when its shard ends earlier (collectives must match across ranks).
constexpr int kBoundExchanges = 4;
constexpr uint64_t kBoundExchangeRows[kBoundExchanges] = {1ULL << 16, 1ULL << 19, 1ULL << 22, 1ULL << 25};

__global__ void fill_sentinel_u32_kernel(uint64_t *ids_out, uint32_t *key_out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { ids_out[i] = UINT64_MAX; key_out[i] = UINT32_MAX; }
}

// ---- exact selection -------------------------------------------------------------------------------
// KeyFn: struct with  __device__ void load_query(uint32_t q)  (all threads of the CTA call it),
// __device__ uint32_t key(uint64_t row) const  (smaller is better), static constexpr int kKeyBits (8 or 32:
// how many low bits of the key are significant) and kInvalidKey (rows with this key never match).
struct ExactScratch {
    unsigned long long hist[256];    // keys 0..255
    unsigned long long digit[256];   // radix pass histogram
    unsigned int out_count;
    unsigned int pad;
};

template <typename KeyFn>
__global__ void __launch_bounds__(256)
exact_select_kernel(KeyFn fn, const uint64_t *__restrict__ ids, uint64_t id_base, uint64_t N,
                    const uint32_t *__restrict__ flags, uint32_t nq, uint32_t k, uint32_t key_flip, ExactScratch *scr,
                    uint64_t *out_id, uint32_t *out_key, uint64_t *out_row, int emit_rows, uint64_t *ids_out, uint32_t *keys_out,
                    unsigned long long *n_selected) {
    namespace cg = cooperative_groups;
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
        if (gtid == 0 && n_selected) atomicAdd(n_selected, 1ULL);
        fn.load_query(q);
        // ---- 1. k-th smallest key: MSB-first 8-bit radix select over KeyFn::kKeyBits bits
        if (gtid == 0) scr->out_count = 0;
        uint32_t kprefix = 0; uint64_t need = k; bool have_k = true;
        for (int shift = KeyFn::kKeyBits - 8; shift >= 0; shift -= 8) {
            if (gtid < 256) scr->hist[gtid] = 0;
            s_hist[threadIdx.x] = 0;
            grid.sync();
            const uint32_t hi_mask = shift + 8 >= 32 ? 0u : ~0u << (shift + 8);
            for (uint64_t r = gtid; r < N; r += gsize) {
                uint32_t key = fn.key(r);
                if ((key & hi_mask) != kprefix) continue;
                atomicAdd(&s_hist[(key >> shift) & 255], 1ULL);
            }
            __syncthreads();
            if (s_hist[threadIdx.x]) atomicAdd(&scr->hist[threadIdx.x], s_hist[threadIdx.x]);
            grid.sync();
            uint64_t cum = 0; uint32_t dig = 256;
            for (uint32_t b = 0; b < 256; ++b) {
                uint64_t h = scr->hist[b];
                if (cum + h >= need) { dig = b; break; }
                cum += h;
            }
            if (dig == 256) { have_k = false; dig = 255; cum = 0; }  // fewer than k rows in total
            need -= cum;
            kprefix |= dig << shift;
            grid.sync();  // everyone has read hist[] before it is cleared again
        }
        const uint32_t kstar = have_k ? kprefix : 0xFFFFFFFFu;
        uint64_t idstar = UINT64_MAX;  // fewer than k rows in total: take everything
        if (have_k) {
            // ---- 2. radix select of the need-th smallest id among rows with key == k*
            uint64_t prefix = 0; uint64_t want = need;
            for (int shift = 56; shift >= 0; shift -= 8) {
                if (gtid < 256) scr->digit[gtid] = 0;
                s_hist[threadIdx.x] = 0;
                grid.sync();
                const uint64_t hi_mask = shift == 56 ? 0 : ~0ULL << (shift + 8);
                for (uint64_t r = gtid; r < N; r += gsize) {
                    if (fn.key(r) != kstar) continue;
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
            uint32_t d = fn.key(r);
            if (d > kstar || d == KeyFn::kInvalidKey) continue;
            uint64_t id = ids ? ids[r] : id_base + r;
            if (d == kstar && id > idstar) continue;
            unsigned int pos = atomicAdd(&scr->out_count, 1u);
            if (pos < k) { out_id[pos] = id; out_key[pos] = d; out_row[pos] = r; }
        }
        grid.sync();
        // ---- 4. CTA 0 orders the winners (k <= 2048: rank by counting, O(k^2) on a tiny set)
        if (blockIdx.x == 0) {
            uint32_t m = min(scr->out_count, k);
            for (uint32_t i = threadIdx.x; i < k; i += blockDim.x)
                if (i >= m) { ids_out[(size_t)q * k + i] = UINT64_MAX; keys_out[(size_t)q * k + i] = UINT32_MAX; }
            for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                uint64_t id = out_id[i]; uint32_t d = out_key[i];
                uint32_t rank = 0;
                for (uint32_t j = 0; j < m; ++j) {
                    uint64_t idj = out_id[j]; uint32_t dj = out_key[j];
                    rank += (dj < d || (dj == d && (idj < id || (idj == id && j < i))));
                }
                ids_out[(size_t)q * k + rank] = emit_rows ? out_row[i] : id;
                keys_out[(size_t)q * k + rank] = KeyFn::report(d, key_flip);
            }
        }
        grid.sync();
    }
}

template <typename KeyFn>
static int exact_select_fallback(ucfp_lane *ctx, ucfp_corpus *c, int occ, KeyFn fn, const uint32_t *flags, uint32_t nq, uint32_t k,
                                 uint32_t key_flip, uint64_t *ids_out, uint32_t *keys_out, int emit_rows = 0) {
    size_t scratch = sizeof(ExactScratch) + (2 * sizeof(uint64_t) + sizeof(uint32_t)) * (size_t)k + 64;
    UCFP_TRY(ctx->misc.reserve(scratch));
    ExactScratch *scr = ctx->misc.as<ExactScratch>();
    uint64_t *out_id = reinterpret_cast<uint64_t *>(scr + 1);
    uint64_t *out_row = out_id + k;
    uint32_t *out_key = reinterpret_cast<uint32_t *>(out_row + k);
    if (occ < 1) occ = 1;   // measured once per context by the *_device_init functions
    const uint64_t *ids = c->id_mode == 1 ? c->ids : nullptr;
    uint64_t id_base = c->id_base, N = c->size;
    unsigned long long *n_selected = ctx->stats.ptr ? ctx->stats.as<unsigned long long>() + 2 : nullptr;   // diagnostics: queries that got this far
    void *args[] = {&fn, &ids, &id_base, &N, &flags, &nq, &k, &key_flip, &scr, &out_id, &out_key, &out_row, &emit_rows, &ids_out, &keys_out, &n_selected};
    UCFP_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)exact_select_kernel<KeyFn>, dim3(ctx->sm_count * occ), dim3(256), args,
                                              0, ctx->stream));
    count_launch(ctx);
    return UCFP_OK;
}
