// Runs the C++ host mirror (include/ucfp/host.hpp) on a real GPU: the reference's own index tests
// (src/index/embedded/mod.rs:523-589) and one image batch, printed so that the Python test can compare
// the hashes with the oracle.
#include <cstdio>
#include <vector>

#include "ucfp/host.hpp"

static ucfp::Record rec(uint32_t tenant, uint64_t rid, std::vector<float> e) {
    ucfp::Record r;
    r.tenant_id = tenant; r.record_id = rid; r.algorithm = "test"; r.embedding = std::move(e);
    return r;
}

int main() {
    try {
        ucfp::Gpu gpu(0);
        {   // upsert_and_knn_round_trip
            ucfp::GpuIndexBackend db(gpu, 1024);
            db.upsert({rec(1, 100, {1.0f, 0.0f, 0.0f}), rec(1, 200, {0.0f, 1.0f, 0.0f}), rec(1, 300, {0.7f, 0.7f, 0.0f})});
            auto hits = db.knn(1, {0.6f, 0.6f, 0.0f}, 2);
            if (hits.size() != 2 || hits[0].record_id != 300 || !(hits[0].score > hits[1].score)) { std::puts("FAIL knn_round_trip"); return 1; }
            for (auto &h : hits) if (h.tenant_id != 1 || h.source != ucfp::HitSource::Vector) { std::puts("FAIL hit fields"); return 1; }
            // knn_ignores_other_tenants
            db.upsert({rec(2, 1, {1.0f, 0.0f, 0.0f})});
            if (db.knn(2, {1.0f, 0.0f, 0.0f}, 10).size() != 1) { std::puts("FAIL tenants"); return 1; }
            if (!db.knn(1, {}, 10).empty() || !db.knn(1, {1.0f, 0.0f, 0.0f}, 0).empty() || !db.knn(9, {1.0f, 0.0f, 0.0f}, 3).empty()) { std::puts("FAIL empty cases"); return 1; }
            ucfp::Query q; q.tenant_id = 1; q.k = 1; q.vector = std::vector<float>{0.0f, 1.0f, 0.0f};
            auto top = ucfp::Matcher(db).search(q);
            if (top.size() != 1 || top[0].record_id != 200) { std::puts("FAIL matcher"); return 1; }
        }
        {   // one multi bundle on the reference ramp image (src/server/tests.rs:227-235), 64 x 64
            const uint32_t w = 64, h = 64;
            std::vector<uint8_t> px(3 * w * h), exact(32, 0xAB);
            for (uint32_t y = 0; y < h; ++y) for (uint32_t x = 0; x < w; ++x) { px[3 * (y * w + x)] = x % 256; px[3 * (y * w + x) + 1] = y % 256; px[3 * (y * w + x) + 2] = 128; }
            auto recs = ucfp::image::fingerprint_batch_rgb(gpu, {ucfp::image::DecodedRgb{px.data(), w, h, exact.data()}}, UCFP_ALGO_MULTI, 7, {42});
            if (recs.size() != 1 || recs[0].fingerprint.size() != 536 || recs[0].algorithm != "imgfprint-multihash-v1") { std::puts("FAIL bundle"); return 1; }
            std::printf("BUNDLE ");
            for (uint8_t b : recs[0].fingerprint) std::printf("%02x", b);
            std::printf("\n");
            try {
                ucfp::image::fingerprint_batch_rgb(gpu, {ucfp::image::DecodedRgb{px.data(), 2, 2, exact.data()}}, UCFP_ALGO_MULTI, 7, {43});
                std::puts("FAIL tiny image accepted"); return 1;
            } catch (const ucfp::Error &e) { if (e.kind != "Modality") { std::puts("FAIL error kind"); return 1; } }
        }
        std::puts("OK");
        return 0;
    } catch (const std::exception &e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 2;
    }
}
