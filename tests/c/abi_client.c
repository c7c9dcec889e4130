/* A plain C11 client of libucfp_cuda.so, run on the GPU by tests/test_host_layer_gpu.py: what the reference's Rust host would
 * do through its FFI crate, written against nothing but include/ucfp_cuda.h.  It replays the reference's own index tests
 * (src/index/embedded/mod.rs:523-589: upsert_and_knn_round_trip, knn_ignores_other_tenants -- one corpus per tenant here),
 * exercises insert-or-replace / delete (src/index/mod.rs:20-25), the batcher, and hashes the reference's ramp image
 * (src/server/tests.rs:227-235), printing the 51 hash words for the Python side to compare with the oracle. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ucfp_cuda.h"

#define CHECK(expr) do { int _rc = (expr); if (_rc != UCFP_OK) { printf("FAIL %s -> %d: %s\n", #expr, _rc, ucfp_last_error()); return 1; } } while (0)
#define EXPECT(cond) do { if (!(cond)) { printf("FAIL %s (line %d)\n", #cond, __LINE__); return 1; } } while (0)

int main(void) {
    ucfp_ctx *ctx = NULL;
    CHECK(ucfp_init(0, &ctx));
    EXPECT(ucfp_abi_version() == UCFP_ABI_VERSION);

    /* upsert_and_knn_round_trip: q = [.6,.6,0] over {[1,0,0] -> 100, [0,1,0] -> 200, [.7,.7,0] -> 300}, k = 2 -> first = 300 */
    ucfp_corpus *t1 = NULL, *t2 = NULL;
    CHECK(ucfp_corpus_create(ctx, UCFP_KIND_COSINE, 3, 2, &t1));   /* capacity 2: the third row makes upsert grow the corpus */
    const uint64_t ids[3] = {100, 200, 300};
    const float rows[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, .7f, .7f, 0.f};
    uint64_t n_rep = 99;
    CHECK(ucfp_corpus_upsert(t1, ids, rows, 3, &n_rep));
    EXPECT(n_rep == 0 && ucfp_corpus_size(t1) == 3 && ucfp_corpus_capacity(t1) >= 3);
    const float q[3] = {.6f, .6f, 0.f};
    uint64_t hit_ids[4]; float hit_scores[4];
    CHECK(ucfp_scan_cosine(t1, q, 1, 2, hit_ids, hit_scores));
    EXPECT(hit_ids[0] == 300 && hit_scores[0] > hit_scores[1]);
    /* knn_ignores_other_tenants */
    CHECK(ucfp_corpus_create(ctx, UCFP_KIND_COSINE, 3, 8, &t2));
    const uint64_t id2 = 1; const float row2[3] = {1.f, 0.f, 0.f};
    CHECK(ucfp_corpus_upsert(t2, &id2, row2, 1, NULL));
    CHECK(ucfp_scan_cosine(t2, row2, 1, 3, hit_ids, hit_scores));
    EXPECT(hit_ids[0] == 1 && hit_ids[1] == UCFP_ID_NONE && hit_ids[2] == UCFP_ID_NONE);
    /* insert-or-replace: record 300 becomes orthogonal to q; delete is idempotent */
    const uint64_t rid = 300; const float rrow[3] = {0.f, 0.f, 1.f};
    CHECK(ucfp_corpus_upsert(t1, &rid, rrow, 1, &n_rep));
    EXPECT(n_rep == 1 && ucfp_corpus_size(t1) == 3);
    CHECK(ucfp_scan_cosine(t1, q, 1, 3, hit_ids, hit_scores));
    EXPECT(hit_ids[2] == 300 && hit_scores[2] == 0.f);
    const uint64_t gone[2] = {100, 424242};
    uint64_t n_rem = 99;
    CHECK(ucfp_corpus_delete(t1, gone, 2, &n_rem));
    EXPECT(n_rem == 1 && ucfp_corpus_size(t1) == 2);
    CHECK(ucfp_corpus_delete(t1, gone, 2, &n_rem));
    EXPECT(n_rem == 0);
    CHECK(ucfp_scan_cosine(t1, q, 1, 3, hit_ids, hit_scores));
    EXPECT(hit_ids[0] == 200 && hit_ids[1] == 300 && hit_ids[2] == UCFP_ID_NONE);
    /* the batcher answers like the direct scan */
    ucfp_batcher *b = NULL;
    CHECK(ucfp_batcher_create(t1, 64, 100, &b));
    uint64_t bid[2]; float bsc[2];
    CHECK(ucfp_batcher_query(b, q, 2, bid, bsc));
    EXPECT(bid[0] == hit_ids[0] && bid[1] == hit_ids[1] && bsc[0] == hit_scores[0]);
    ucfp_batcher_destroy(b);
    /* error behaviour: wrong kind, bad arguments -> negative status + message, never a crash */
    uint32_t dist[2];
    EXPECT(ucfp_scan_hamming(t1, ids, 1, 2, hit_ids, dist) == UCFP_E_STATE && strlen(ucfp_last_error()) > 0);
    EXPECT(ucfp_corpus_create(ctx, 99, 0, 10, &t2) == UCFP_E_INVALID);
    ucfp_corpus_destroy(t1);

    /* one multi bundle on the reference ramp image, 64 x 64 */
    enum { W = 64, H = 64 };
    static uint8_t px[3 * W * H];
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) { px[3 * (y * W + x)] = (uint8_t)x; px[3 * (y * W + x) + 1] = (uint8_t)y; px[3 * (y * W + x) + 2] = 128; }
    ucfp_image_desc d = {px, W, H, 3 * W};
    ucfp_image_desc tiny = {px, 2, 2, 6};
    ucfp_image_desc two[2]; two[0] = d; two[1] = tiny;
    ucfp_image_hashes hh[2]; int32_t st[2];
    CHECK(ucfp_image_hash_batch(ctx, two, 2, UCFP_ALGO_MULTI, hh, st));
    EXPECT(st[0] == UCFP_OK && st[1] == UCFP_E_INVALID);            /* one bad image does not fail the batch */
    const uint64_t *wds = (const uint64_t *)&hh[0];
    printf("WORDS");
    for (int i = 0; i < 51; ++i) printf(" %016llx", (unsigned long long)wds[i]);
    printf("\n");
    ucfp_destroy(ctx);
    puts("OK");
    return 0;
}
