/*
 * ucfp_oracle.c -- CPU restatement of the UCFP fingerprint hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ucfp_b200/, include/)
 * may include, link or call this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs load it, as the checker
 * and as the timed CPU baseline.
 *
 * What it restates (citations relative to /root/reference):
 *   - cosine top-k  : src/index/embedded/mod.rs:268-360 (knn), :454-472
 *                     (dot_product, 8 f32 lanes), :475 (l2_norm), :484-495
 *                     (insert_topk).  This arithmetic is fully in-tree, so
 *                     cosine parity is PINNED by the reference's own
 *                     known-answer tests (embedded/mod.rs:523-589,
 *                     server/tests.rs:53-113), ported in tests/.
 *   - Hamming top-k : NOT in the reference (no server-side compare exists;
 *                     docs tell clients to run BIT_COUNT(phash ^ ?) ORDER BY d,
 *                     web/src/lib/docs/examples.md:72-76).  Defined by
 *                     docs/HASH_SPEC.md section 6: popcount(q ^ c), k smallest,
 *                     total order (distance asc, record_id asc).
 *   - Jaccard top-k : NOT in the reference.  MinHashSig<128> layout from
 *                     src/modality/text.rs:200-204 and server/tests.rs:1153-1162;
 *                     estimator = equal slots / 128
 *                     (web/.../MinHashSlotHeatmap.svelte:86-92).  Total order
 *                     (matches desc, record_id asc).
 *   - image hashes  : the arithmetic lives in the third-party crate
 *                     imgfprint 0.4.1 (Cargo.lock:1863), whose source is not
 *                     under /root/reference; call sites src/modality/image.rs:
 *                     68-70,175-179.  The reference's tests pin no hash bit
 *                     (server/tests.rs:239-263,456-532,1167-1208 check tags and
 *                     the 536-byte size only).  PARITY UNPINNED for image hash
 *                     bits: this file implements docs/HASH_SPEC.md, which
 *                     restates the published `image` 0.25 Triangle resize
 *                     (the filter the reference itself uses in
 *                     src/modality/image.rs:310-319) and the classic
 *                     AHash/DHash/PHash definitions.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -ffp-contract=off; the
 * reference ships with target-cpu=x86-64-v3, .cargo/config.toml:15-16, and Rust
 * never contracts a*b+c into an FMA).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dct_table.h"

#define UCFP_ORACLE_API __attribute__((visibility("default")))
#define ID_NONE UINT64_MAX

/* ------------------------------------------------------------------------ */
/* Deterministic counter-based PRNG shared (by definition, not by code) with  */
/* the device generators: docs/HASH_SPEC.md section 8.                        */
/* ------------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

UCFP_ORACLE_API uint64_t ucfp_oracle_splitmix64(uint64_t seed, uint64_t index) {
    return mix64(seed ^ ((index + 1) * 0x9E3779B97F4A7C15ULL));
}

UCFP_ORACLE_API void ucfp_oracle_fill_u64(uint64_t *out, size_t n, uint64_t seed, uint64_t start) {
    for (size_t i = 0; i < n; ++i) out[i] = ucfp_oracle_splitmix64(seed, start + i);
}

/* ------------------------------------------------------------------------ */
/* Small thread helper: static row partition, mirrors rayon fold/reduce       */
/* (embedded/mod.rs:324-340) without work stealing.                           */
/* ------------------------------------------------------------------------ */
typedef void (*range_fn)(void *arg, int tid, size_t lo, size_t hi);
typedef struct { range_fn fn; void *arg; int tid; size_t lo, hi; } job_t;
static void *job_tramp(void *p) { job_t *j = (job_t *)p; j->fn(j->arg, j->tid, j->lo, j->hi); return NULL; }

static void parallel_ranges(size_t n, int threads, range_fn fn, void *arg) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    if (threads == 1) { fn(arg, 0, 0, n); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t].fn = fn; jobs[t].arg = arg; jobs[t].tid = t;
        jobs[t].lo = n * (size_t)t / threads; jobs[t].hi = n * (size_t)(t + 1) / threads;
        pthread_create(&th[t], NULL, job_tramp, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* the same stream, filled by several host threads (large synthetic corpora of the full-size parity checks) */
typedef struct { uint64_t *out; uint64_t seed, start; } fill_job_t;
static void fill_range(void *arg, int tid, size_t lo, size_t hi) {
    (void)tid;
    fill_job_t *f = (fill_job_t *)arg;
    for (size_t i = lo; i < hi; ++i) f->out[i] = ucfp_oracle_splitmix64(f->seed, f->start + i);
}
UCFP_ORACLE_API void ucfp_oracle_fill_u64_mt(uint64_t *out, size_t n, uint64_t seed, uint64_t start, int threads) {
    fill_job_t f = {out, seed, start};
    parallel_ranges(n, threads, fill_range, &f);
}

/* ------------------------------------------------------------------------ */
/* Integer-keyed top-k with a total order (key asc, id asc).                  */
/* Hamming uses key = distance; Jaccard uses key = 128 - matches.             */
/* ------------------------------------------------------------------------ */
typedef struct { uint32_t key; uint64_t id; } ikey_t;

static inline int ikey_less(uint32_t ka, uint64_t ia, uint32_t kb, uint64_t ib) {
    return ka < kb || (ka == kb && ia < ib);
}

/* buf holds *len <= k entries sorted ascending */
static inline void itopk_insert(ikey_t *buf, size_t *len, size_t k, uint32_t key, uint64_t id) {
    size_t n = *len;
    if (n == k) {
        if (!ikey_less(key, id, buf[n - 1].key, buf[n - 1].id)) return;
        n--;
    }
    size_t pos = n;
    while (pos > 0 && ikey_less(key, id, buf[pos - 1].key, buf[pos - 1].id)) { buf[pos] = buf[pos - 1]; pos--; }
    buf[pos].key = key; buf[pos].id = id;
    *len = n + 1;
}

typedef struct {
    const uint64_t *rows; const uint64_t *ids; uint64_t id_base; size_t n;
    const uint64_t *q; size_t nq; size_t k; int kind; /* 0 hamming, 1 jaccard */
    ikey_t *partial; size_t *partial_len; /* [threads][nq][k] */
} iscan_t;

static inline uint32_t jaccard_matches(const uint64_t *a, const uint64_t *b) {
    uint32_t m = 0;
    for (int i = 0; i < 128; ++i) m += (a[i] == b[i]);
    return m;
}

static void iscan_range(void *arg, int tid, size_t lo, size_t hi) {
    iscan_t *s = (iscan_t *)arg;
    /* block the rows so that every query sees a cache-resident slab */
    const size_t slab = s->kind == 0 ? 4096 : 64;
    for (size_t b = lo; b < hi; b += slab) {
        size_t e = b + slab < hi ? b + slab : hi;
        for (size_t qi = 0; qi < s->nq; ++qi) {
            ikey_t *buf = s->partial + ((size_t)tid * s->nq + qi) * s->k;
            size_t *len = s->partial_len + (size_t)tid * s->nq + qi;
            if (s->kind == 0) {
                const uint64_t qc = s->q[qi];
                /* thr = worst key currently kept (65 while the buffer is not full): rows with d > thr cannot enter */
                uint32_t thr = *len == s->k ? buf[s->k - 1].key : 65u;
                for (size_t r = b; r < e; ++r) {
                    uint32_t d = (uint32_t)__builtin_popcountll(qc ^ s->rows[r]);
                    if (__builtin_expect(d > thr, 1)) continue;
                    itopk_insert(buf, len, s->k, d, s->ids ? s->ids[r] : s->id_base + r);
                    thr = *len == s->k ? buf[s->k - 1].key : 65u;
                }
            } else {
                const uint64_t *qs = s->q + qi * 128;
                for (size_t r = b; r < e; ++r) {
                    uint32_t key = 128u - jaccard_matches(qs, s->rows + r * 128);
                    if (*len == s->k && key > buf[s->k - 1].key) continue;
                    itopk_insert(buf, len, s->k, key, s->ids ? s->ids[r] : s->id_base + r);
                }
            }
        }
    }
}

static void iscan_topk(int kind, const uint64_t *rows, const uint64_t *ids, uint64_t id_base, size_t n,
                       const uint64_t *q, size_t nq, size_t k, uint64_t *ids_out, uint32_t *key_out, int threads) {
    if (nq == 0 || k == 0) return;
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    iscan_t s = { rows, ids, id_base, n, q, nq, k, kind, NULL, NULL };
    s.partial = (ikey_t *)malloc(sizeof(ikey_t) * (size_t)threads * nq * k);
    s.partial_len = (size_t *)calloc((size_t)threads * nq, sizeof(size_t));
    parallel_ranges(n, threads, iscan_range, &s);
    for (size_t qi = 0; qi < nq; ++qi) {
        ikey_t *dst = s.partial + qi * k; size_t *dlen = s.partial_len + qi; /* thread 0 buffer */
        for (int t = 1; t < threads; ++t) {
            ikey_t *src = s.partial + ((size_t)t * nq + qi) * k; size_t slen = s.partial_len[(size_t)t * nq + qi];
            for (size_t i = 0; i < slen; ++i) itopk_insert(dst, dlen, k, src[i].key, src[i].id);
        }
        for (size_t i = 0; i < k; ++i) {
            if (i < *dlen) { ids_out[qi * k + i] = dst[i].id; key_out[qi * k + i] = kind == 0 ? dst[i].key : 128u - dst[i].key; }
            else { ids_out[qi * k + i] = ID_NONE; key_out[qi * k + i] = UINT32_MAX; }
        }
    }
    free(s.partial); free(s.partial_len);
}

/* Hamming top-k: docs/HASH_SPEC.md section 6.  ids may be NULL (id = id_base + row). */
UCFP_ORACLE_API void ucfp_oracle_hamming_topk(const uint64_t *codes, const uint64_t *ids, uint64_t id_base, size_t n,
                                              const uint64_t *queries, size_t nq, size_t k,
                                              uint64_t *ids_out, uint32_t *dist_out, int threads) {
    iscan_topk(0, codes, ids, id_base, n, queries, nq, k, ids_out, dist_out, threads);
}

/* MinHash-128 Jaccard top-k: docs/HASH_SPEC.md section 7.  sigs = n x 128 u64 slot payloads. */
UCFP_ORACLE_API void ucfp_oracle_jaccard_topk(const uint64_t *sigs, const uint64_t *ids, uint64_t id_base, size_t n,
                                              const uint64_t *queries, size_t nq, size_t k,
                                              uint64_t *ids_out, uint32_t *matches_out, int threads) {
    iscan_topk(1, sigs, ids, id_base, n, queries, nq, k, ids_out, matches_out, threads);
}

/* ------------------------------------------------------------------------ */
/* Cosine k-NN: restatement of src/index/embedded/mod.rs.                     */
/* ------------------------------------------------------------------------ */

/* embedded/mod.rs:454-472 -- eight independent f32 accumulators over chunks of
 * eight, summed left to right, then the scalar remainder. */
UCFP_ORACLE_API float ucfp_oracle_dot_product(const float *a, const float *b, size_t n) {
    float accs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t chunks = n / 8;
    for (size_t c = 0; c < chunks; ++c)
        for (int j = 0; j < 8; ++j) accs[j] += a[c * 8 + j] * b[c * 8 + j];
    float sum = 0.0f; /* Iterator::sum starts from 0.0 and adds accs[0..8] in order */
    for (int j = 0; j < 8; ++j) sum += accs[j];
    for (size_t i = chunks * 8; i < n; ++i) sum += a[i] * b[i];
    return sum;
}

/* embedded/mod.rs:475 */
UCFP_ORACLE_API float ucfp_oracle_l2_norm(const float *v, size_t n) { return sqrtf(ucfp_oracle_dot_product(v, v, n)); }

typedef struct { uint64_t id; float score; } fhit_t;

/* embedded/mod.rs:484-495 -- sorted-descending buffer; insertion point = first
 * element whose score is NOT > the new score (partition_point(|s| s > score)),
 * so a new tie lands in front of older ties; a full buffer admits only
 * score > worst. */
static inline void ref_insert_topk(fhit_t *local, size_t *len, uint64_t rid, float score, size_t k) {
    size_t n = *len;
    if (n < k) {
        size_t pos = 0; while (pos < n && local[pos].score > score) pos++;
        memmove(local + pos + 1, local + pos, (n - pos) * sizeof(fhit_t));
        local[pos].id = rid; local[pos].score = score; *len = n + 1;
    } else if (n > 0 && score > local[n - 1].score) {
        size_t pos = 0; while (pos < n && local[pos].score > score) pos++;
        memmove(local + pos + 1, local + pos, (n - 1 - pos) * sizeof(fhit_t));
        local[pos].id = rid; local[pos].score = score;
    }
}

/* total-order variant used to compare with the GPU: (score desc, id asc) */
static inline int fhit_before(float sa, uint64_t ia, float sb, uint64_t ib) { return sa > sb || (sa == sb && ia < ib); }
static inline void tot_insert_topk(fhit_t *local, size_t *len, uint64_t rid, float score, size_t k) {
    size_t n = *len;
    if (n == k) { if (!fhit_before(score, rid, local[n - 1].score, local[n - 1].id)) return; n--; }
    size_t pos = n;
    while (pos > 0 && fhit_before(score, rid, local[pos - 1].score, local[pos - 1].id)) { local[pos] = local[pos - 1]; pos--; }
    local[pos].id = rid; local[pos].score = score; *len = n + 1;
}

typedef struct {
    const float *rows; const uint64_t *ids; uint64_t id_base; size_t n, dim;
    const float *q; size_t nq, k; int mode; const float *qnorm;
    fhit_t *partial; size_t *partial_len;
} cscan_t;

static void cscan_range(void *arg, int tid, size_t lo, size_t hi) {
    cscan_t *s = (cscan_t *)arg;
    const size_t slab = 256;
    for (size_t b = lo; b < hi; b += slab) {
        size_t e = b + slab < hi ? b + slab : hi;
        for (size_t qi = 0; qi < s->nq; ++qi) {
            float q_norm = s->qnorm[qi];
            if (q_norm == 0.0f) continue;                    /* embedded/mod.rs:283-286 */
            fhit_t *buf = s->partial + ((size_t)tid * s->nq + qi) * s->k;
            size_t *len = s->partial_len + (size_t)tid * s->nq + qi;
            const float *qv = s->q + qi * s->dim;
            for (size_t r = b; r < e; ++r) {
                const float *v = s->rows + r * s->dim;
                float v_norm = ucfp_oracle_l2_norm(v, s->dim);  /* :327 */
                if (v_norm == 0.0f) continue;                   /* :328-330 */
                float score = ucfp_oracle_dot_product(qv, v, s->dim) / (q_norm * v_norm); /* :331 */
                uint64_t id = s->ids ? s->ids[r] : s->id_base + r;
                if (s->mode == 0) ref_insert_topk(buf, len, id, score, s->k);
                else tot_insert_topk(buf, len, id, score, s->k);
            }
        }
    }
}

/*
 * mode 0: the reference's own insert_topk tie behaviour, rows visited in index
 *         order per thread, thread partials reduced in thread order, then the
 *         stable sort of embedded/mod.rs:342 (a no-op on a sorted buffer).
 *         With threads == 1 this is exactly the sequential reference result.
 * mode 1: total order (score desc, id asc) -- what the GPU path guarantees.
 * count_out[q] = number of valid hits (rows with zero norm are skipped, so it
 * may be < min(k, n)); remaining slots get id = UINT64_MAX, score = -inf.
 * Empty query (dim == 0), k == 0 or zero-norm query -> count 0 (:275, :284).
 */
UCFP_ORACLE_API void ucfp_oracle_cosine_topk(const float *rows, const uint64_t *ids, uint64_t id_base, size_t n, size_t dim,
                                             const float *queries, size_t nq, size_t k, int mode,
                                             uint64_t *ids_out, float *score_out, uint32_t *count_out, int threads) {
    if (nq == 0) return;
    for (size_t qi = 0; qi < nq; ++qi) if (count_out) count_out[qi] = 0;
    if (k == 0 || dim == 0) return;
    for (size_t i = 0; i < nq * k; ++i) { ids_out[i] = ID_NONE; score_out[i] = -INFINITY; }
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    float *qnorm = (float *)malloc(sizeof(float) * nq);
    for (size_t qi = 0; qi < nq; ++qi) qnorm[qi] = ucfp_oracle_l2_norm(queries + qi * dim, dim);
    cscan_t s = { rows, ids, id_base, n, dim, queries, nq, k, mode, qnorm, NULL, NULL };
    s.partial = (fhit_t *)malloc(sizeof(fhit_t) * (size_t)threads * nq * k);
    s.partial_len = (size_t *)calloc((size_t)threads * nq, sizeof(size_t));
    parallel_ranges(n, threads, cscan_range, &s);
    for (size_t qi = 0; qi < nq; ++qi) {
        fhit_t *dst = s.partial + qi * k; size_t *dlen = s.partial_len + qi;
        for (int t = 1; t < threads; ++t) {
            fhit_t *src = s.partial + ((size_t)t * nq + qi) * k; size_t slen = s.partial_len[(size_t)t * nq + qi];
            for (size_t i = 0; i < slen; ++i) {
                if (mode == 0) ref_insert_topk(dst, dlen, src[i].id, src[i].score, k);   /* reduce, :335-340 */
                else tot_insert_topk(dst, dlen, src[i].id, src[i].score, k);
            }
        }
        for (size_t i = 0; i < *dlen; ++i) { ids_out[qi * k + i] = dst[i].id; score_out[qi * k + i] = dst[i].score; }
        if (count_out) count_out[qi] = (uint32_t)*dlen;
    }
    free(s.partial); free(s.partial_len); free(qnorm);
}

/* ------------------------------------------------------------------------ */
/* Image hashing: docs/HASH_SPEC.md sections 1-5.                              */
/* ------------------------------------------------------------------------ */

/* `image` 0.25 rgb_to_luma for u8: (2126 R + 7152 G + 722 B) / 10000, integer
 * truncating division (spec section 1). */
UCFP_ORACLE_API void ucfp_oracle_gray(const uint8_t *rgb, int w, int h, size_t stride, uint8_t *out) {
    for (int y = 0; y < h; ++y) {
        const uint8_t *row = rgb + (size_t)y * stride;
        for (int x = 0; x < w; ++x) {
            uint32_t l = 2126u * row[3 * x] + 7152u * row[3 * x + 1] + 722u * row[3 * x + 2];
            out[(size_t)y * w + x] = (uint8_t)(l / 10000u);
        }
    }
}

typedef struct { int left, n; float *w; } taps_t;

/* Weight construction of `image` 0.25 imageops::sample::{vertical,horizontal}_sample
 * with the Triangle kernel (support 1.0), all arithmetic in f32 (spec section 2). */
static taps_t *make_taps(int src, int dst) {
    taps_t *t = (taps_t *)malloc(sizeof(taps_t) * dst);
    float ratio = (float)src / (float)dst;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float support = 1.0f * sratio;
    for (int o = 0; o < dst; ++o) {
        float in = ((float)o + 0.5f) * ratio;
        int64_t left = (int64_t)floorf(in - support);
        if (left < 0) left = 0;
        if (left > src - 1) left = src - 1;
        int64_t right = (int64_t)ceilf(in + support);
        if (right < left + 1) right = left + 1;
        if (right > src) right = src;
        in = in - 0.5f;
        int n = (int)(right - left);
        float *w = (float *)malloc(sizeof(float) * n);
        float sum = 0.0f;
        for (int i = 0; i < n; ++i) {
            float x = ((float)(left + i) - in) / sratio;
            float a = fabsf(x);
            w[i] = a < 1.0f ? 1.0f - a : 0.0f;
            sum += w[i];
        }
        for (int i = 0; i < n; ++i) w[i] /= sum;
        t[o].left = (int)left; t[o].n = n; t[o].w = w;
    }
    return t;
}
static void free_taps(taps_t *t, int dst) { for (int o = 0; o < dst; ++o) free(t[o].w); free(t); }

/* Export of the tap table so tests can compare the library's host-side table builder. */
UCFP_ORACLE_API int ucfp_oracle_triangle_taps(int src, int dst, int o, int *left, float *w, int cap) {
    taps_t *t = make_taps(src, dst);
    int n = t[o].n; *left = t[o].left;
    for (int i = 0; i < n && i < cap; ++i) w[i] = t[o].w[i];
    free_taps(t, dst);
    return n;
}

/* resize(gray region, nw, nh, Triangle): vertical pass into f32, then
 * horizontal pass, clamp to [0,255], round half away from zero (spec section 2). */
UCFP_ORACLE_API void ucfp_oracle_resize_triangle(const uint8_t *gray, int w, int h, size_t stride,
                                                 int nw, int nh, uint8_t *out) {
    if (nw == w && nh == h) {                      /* imageops::resize copies when dimensions match */
        for (int y = 0; y < h; ++y) memcpy(out + (size_t)y * nw, gray + (size_t)y * stride, (size_t)w);
        return;
    }
    taps_t *tv = make_taps(h, nh), *th = make_taps(w, nw);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)w * nh);
    for (int oy = 0; oy < nh; ++oy)
        for (int x = 0; x < w; ++x) {
            float t = 0.0f;
            for (int i = 0; i < tv[oy].n; ++i) t += (float)gray[(size_t)(tv[oy].left + i) * stride + x] * tv[oy].w[i];
            tmp[(size_t)oy * w + x] = t;
        }
    for (int ox = 0; ox < nw; ++ox)
        for (int y = 0; y < nh; ++y) {
            float t = 0.0f;
            for (int i = 0; i < th[ox].n; ++i) t += tmp[(size_t)y * w + th[ox].left + i] * th[ox].w[i];
            if (t < 0.0f) t = 0.0f;
            if (t > 255.0f) t = 255.0f;
            out[(size_t)y * nw + ox] = (uint8_t)roundf(t);
        }
    free(tmp); free_taps(tv, nh); free_taps(th, nw);
}

/* AHash over an 8x8 grid: bit 8r+c set iff pixel > mean(64 pixels); decided in
 * integers as 64*p > sum, identical to the f32 or the floor-mean compare
 * (spec section 3; src/modality/image.rs:315-318 uses the integer mean). */
UCFP_ORACLE_API uint64_t ucfp_oracle_ahash_bits(const uint8_t g8[64]) {
    uint32_t sum = 0; for (int i = 0; i < 64; ++i) sum += g8[i];
    uint64_t bits = 0;
    for (int i = 0; i < 64; ++i) if (64u * g8[i] > sum) bits |= 1ULL << i;
    return bits;
}

/* DHash over a 9-wide x 8-tall grid: bit 8r+c set iff g[r][c] > g[r][c+1] (spec section 5). */
UCFP_ORACLE_API uint64_t ucfp_oracle_dhash_bits(const uint8_t g98[72]) {
    uint64_t bits = 0;
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) if (g98[r * 9 + c] > g98[r * 9 + c + 1]) bits |= 1ULL << (8 * r + c);
    return bits;
}

/* PHash: un-normalised DCT-II of the 32x32 grid, low 8x8 block, rows first then
 * columns, sequential f32 mul+add; threshold = mean of the two middle values of
 * the 64 sorted coefficients; bit 8u+v set iff D[u][v] > median (spec section 4). */
UCFP_ORACLE_API uint64_t ucfp_oracle_phash_bits(const uint8_t g32[1024], float *coeff_out /* nullable, 64 */) {
    float R[32][8];
    for (int y = 0; y < 32; ++y)
        for (int v = 0; v < 8; ++v) {
            float t = 0.0f;
            for (int x = 0; x < 32; ++x) t += (float)g32[y * 32 + x] * ucfp_oracle_dct_cos[v][x];
            R[y][v] = t;
        }
    float D[64];
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < 8; ++v) {
            float t = 0.0f;
            for (int y = 0; y < 32; ++y) t += ucfp_oracle_dct_cos[u][y] * R[y][v];
            D[u * 8 + v] = t;
        }
    if (coeff_out) memcpy(coeff_out, D, sizeof(D));
    float s[64]; memcpy(s, D, sizeof(D));
    for (int i = 1; i < 64; ++i) { float x = s[i]; int j = i; while (j > 0 && s[j - 1] > x) { s[j] = s[j - 1]; j--; } s[j] = x; }
    float median = (s[31] + s[32]) * 0.5f;
    uint64_t bits = 0;
    for (int i = 0; i < 64; ++i) if (D[i] > median) bits |= 1ULL << i;
    return bits;
}

static void hash_region(const uint8_t *gray, int w, int h, size_t stride, uint64_t *ah, uint64_t *ph, uint64_t *dh) {
    uint8_t g32[1024], g98[72], g8[64];
    ucfp_oracle_resize_triangle(gray, w, h, stride, 32, 32, g32);
    ucfp_oracle_resize_triangle(gray, w, h, stride, 9, 8, g98);
    ucfp_oracle_resize_triangle(gray, w, h, stride, 8, 8, g8);
    *ah = ucfp_oracle_ahash_bits(g8);
    *ph = ucfp_oracle_phash_bits(g32, NULL);
    *dh = ucfp_oracle_dhash_bits(g98);
}

/*
 * Multi bundle for one decoded RGB8 image: out[0..17) = AHash (global, then the
 * 16 blocks 4r+c), out[17..34) = PHash, out[34..51) = DHash -- the order of the
 * 536-byte MultiHashFingerprint after its 32-byte `exact` prefix
 * (web/src/lib/components/charts/AlgorithmView.svelte:30-37).  Block (r, c)
 * covers rows [r*h/4, (r+1)*h/4) and columns [c*w/4, (c+1)*w/4) (spec section 5).
 * Returns 0, or -1 when w or h < 4 (no 4x4 grid exists).
 */
UCFP_ORACLE_API int ucfp_oracle_image_multihash(const uint8_t *rgb, int w, int h, size_t stride, uint64_t out[51]) {
    if (w < 4 || h < 4) return -1;
    uint8_t *gray = (uint8_t *)malloc((size_t)w * h);
    ucfp_oracle_gray(rgb, w, h, stride, gray);
    hash_region(gray, w, h, (size_t)w, &out[0], &out[17], &out[34]);
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            int y0 = (int)((int64_t)r * h / 4), y1 = (int)((int64_t)(r + 1) * h / 4);
            int x0 = (int)((int64_t)c * w / 4), x1 = (int)((int64_t)(c + 1) * w / 4);
            int b = 1 + 4 * r + c;
            hash_region(gray + (size_t)y0 * w + x0, x1 - x0, y1 - y0, (size_t)w, &out[b], &out[17 + b], &out[34 + b]);
        }
    free(gray);
    return 0;
}

typedef struct { const uint8_t *rgb; int w, h; size_t stride, img_stride; uint64_t *out; } ibatch_t;
static void ibatch_range(void *arg, int tid, size_t lo, size_t hi) {
    (void)tid; ibatch_t *s = (ibatch_t *)arg;
    for (size_t i = lo; i < hi; ++i) ucfp_oracle_image_multihash(s->rgb + i * s->img_stride, s->w, s->h, s->stride, s->out + i * 51);
}

/* n same-sized images, img_stride bytes apart; threads = host threads to use. */
UCFP_ORACLE_API void ucfp_oracle_image_multihash_batch(const uint8_t *rgb, size_t n, int w, int h, size_t stride,
                                                       size_t img_stride, uint64_t *out, int threads) {
    ibatch_t s = { rgb, w, h, stride, img_stride, out };
    parallel_ranges(n, threads, ibatch_range, &s);
}

/* Synthetic images (spec section 8) are the little-endian byte view of
 * ucfp_oracle_fill_u64; the reference ramp of benches/end_to_end.rs:77-85 is
 * produced by the tests. */

/* ------------------------------------------------------------------------ */
/* Multi-hash compare (docs/HASH_SPEC.md section 10).  The weights are the    */
/* reference's MultiHashConfigDto (src/server/dto.rs:462-480) with the        */
/* defaults of web/src/lib/docs/api-reference-image.md:51-62; the compare     */
/* itself lives in imgfprint (not readable here) and is fixed by the spec.    */
/* cfg = {ahash_weight, phash_weight, dhash_weight, global_weight,            */
/*        block_weight}; volatile keeps every f32 operation separately        */
/* rounded in the order the spec writes them.                                 */
/* ------------------------------------------------------------------------ */
UCFP_ORACLE_API float ucfp_oracle_multihash_score(const uint64_t *x, const uint64_t *y, const float *cfg, uint32_t block_thr) {
    volatile float s[3];
    for (int a = 0; a < 3; ++a) {
        const uint64_t *hx = x + 17 * a, *hy = y + 17 * a;
        int dg = __builtin_popcountll(hx[0] ^ hy[0]), m = 0;
        for (int i = 1; i <= 16; ++i) m += (uint32_t)__builtin_popcountll(hx[i] ^ hy[i]) <= block_thr;
        volatile float g = (float)(64 - dg) / 64.0f, b = (float)m / 16.0f;
        volatile float t1 = cfg[3] * g, t2 = cfg[4] * b, t3 = t1 + t2, den = cfg[3] + cfg[4];
        s[a] = den == 0.0f ? 0.0f : t3 / den;
    }
    volatile float u0 = cfg[0] * s[0], u1 = cfg[1] * s[1], u2 = cfg[2] * s[2];
    volatile float num = u0 + u1;
    num = num + u2;
    volatile float den = cfg[0] + cfg[1];
    den = den + cfg[2];
    return den == 0.0f ? 0.0f : num / den;
}
