// host.hpp -- header-only C++17 mirror of the reference's Rust interfaces for the accelerated path, sitting
// directly on the C ABI (ucfp_cuda.h).  Same names, argument meaning and error behaviour as the reference:
//   core types    src/core/mod.rs:19-189        Modality, Record, Hit, HitSource, Query
//   errors        src/error.rs:9-61             ucfp::Error{kind, message} thrown where Rust returns Err
//   image         src/modality/image.rs:38-194  ALGORITHM_* tags, fingerprint_*_rgb (decode stays with the host)
//   index         src/index/mod.rs:17-78        GpuIndexBackend::knn (+ hamming_knn / jaccard_knn, new)
//   matcher       src/matcher/mod.rs:140-207    Matcher::search, vector arm
// The reference's toolchain (cargo) is absent from the development image, so this mirror is what a C++ host
// links today; the Rust crate with the same shape is rust/ucfp-cuda/.
#pragma once

#include <cstdint>
#include <cstring>
#include <array>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../ucfp_cuda.h"

namespace ucfp {

enum class Modality { Audio, Image, Text };
enum class HitSource { Vector, Bm25, Filter, Reranker, Fused };

struct Error : std::runtime_error {
    std::string kind;  // "Modality", "Index", "Unsupported", ... (src/error.rs variants)
    Error(std::string k, const std::string &msg) : std::runtime_error(k + ": " + msg), kind(std::move(k)) {}
};

struct Record {  // src/core/mod.rs:34-72
    uint32_t tenant_id = 0;
    uint64_t record_id = 0;
    Modality modality = Modality::Image;
    uint32_t format_version = 1;
    std::string algorithm;
    uint64_t config_hash = 0;
    std::vector<uint8_t> fingerprint;
    std::optional<std::vector<float>> embedding;
    std::optional<std::string> model_id;
    std::vector<uint8_t> metadata;
    std::optional<std::string> text;
};

struct Hit {  // src/core/mod.rs:108-131
    uint32_t tenant_id = 0;
    uint64_t record_id = 0;
    float score = 0;
    HitSource source = HitSource::Vector;
    std::optional<float> vector_score, bm25_score;
    std::optional<uint32_t> vector_rank, bm25_rank;
};

struct Query {  // src/core/mod.rs:153-189
    uint32_t tenant_id = 0;
    Modality modality = Modality::Text;
    size_t k = 10;
    std::optional<std::vector<float>> vector;
    std::vector<std::string> terms;
    uint32_t rrf_k = 60;
    bool explain = false;
    // not in the reference: the two query kinds its stored fingerprints call for
    std::optional<uint64_t> hash;                       // 64-bit perceptual-hash code -> Hamming top-k
    std::string hash_algorithm;                         // algorithm tag of the stored fingerprints to search (image.rs:38-46)
    std::optional<std::array<uint64_t, 128>> signature; // MinHash-128 slots -> Jaccard top-k
};

namespace detail {
inline void check(int rc, const char *kind) {
    if (rc != UCFP_OK) throw Error(rc == UCFP_E_UNSUPPORTED ? "Unsupported" : kind, ucfp_last_error());
}
}  // namespace detail

class Gpu {  // one per (process, GPU)
   public:
    explicit Gpu(int device = 0) { detail::check(ucfp_init(device, &ctx_), "Index"); }
    ~Gpu() { ucfp_destroy(ctx_); }
    Gpu(const Gpu &) = delete;
    Gpu &operator=(const Gpu &) = delete;
    ucfp_ctx *raw() const { return ctx_; }

   private:
    ucfp_ctx *ctx_ = nullptr;
};

namespace image {
constexpr const char *ALGORITHM_MULTIHASH = "imgfprint-multihash-v1";  // src/modality/image.rs:40
constexpr const char *ALGORITHM_PHASH = "imgfprint-phash-v1";
constexpr const char *ALGORITHM_DHASH = "imgfprint-dhash-v1";
constexpr const char *ALGORITHM_AHASH = "imgfprint-ahash-v1";

struct DecodedRgb { const uint8_t *pixels; uint32_t width, height; const uint8_t *exact32; /* BLAKE3 of the encoded bytes (host) */ };

inline std::vector<uint8_t> pack_single(const uint8_t *exact32, const ucfp_hash17 &h) {  // 168-byte ImageFingerprint
    std::vector<uint8_t> b(168);
    std::memcpy(b.data(), exact32, 32);
    std::memcpy(b.data() + 32, &h, 136);
    return b;
}

// Batched fingerprinting after host-side decode.  One Record per image; a failing image throws Error("Modality")
// only for itself when `errors` is null, otherwise its message is stored and the record left empty.
inline std::vector<Record> fingerprint_batch_rgb(const Gpu &gpu, const std::vector<DecodedRgb> &imgs, uint32_t algo_mask,
                                                 uint32_t tenant_id, const std::vector<uint64_t> &record_ids,
                                                 std::vector<std::string> *errors = nullptr) {
    std::vector<ucfp_image_desc> d(imgs.size());
    for (size_t i = 0; i < imgs.size(); ++i) d[i] = ucfp_image_desc{imgs[i].pixels, imgs[i].width, imgs[i].height, 3ull * imgs[i].width};
    std::vector<ucfp_image_hashes> out(imgs.size());
    std::vector<int32_t> status(imgs.size());
    detail::check(ucfp_image_hash_batch(gpu.raw(), d.data(), d.size(), algo_mask, out.data(), status.data()), "Modality");
    std::vector<Record> recs(imgs.size());
    if (errors) errors->assign(imgs.size(), "");
    for (size_t i = 0; i < imgs.size(); ++i) {
        if (status[i] != UCFP_OK) {
            if (!errors) throw Error("Modality", "image hash failed with status " + std::to_string(status[i]));
            (*errors)[i] = "image hash failed with status " + std::to_string(status[i]);
            continue;
        }
        Record &r = recs[i];
        r.tenant_id = tenant_id; r.record_id = record_ids[i]; r.modality = Modality::Image; r.config_hash = 0;
        if (algo_mask == UCFP_ALGO_MULTI) {  // 536 bytes: exact | ahash | phash | dhash (each 168)
            r.algorithm = ALGORITHM_MULTIHASH;
            r.fingerprint.assign(imgs[i].exact32, imgs[i].exact32 + 32);
            for (const ucfp_hash17 *h : {&out[i].ahash, &out[i].phash, &out[i].dhash}) {
                auto s = pack_single(imgs[i].exact32, *h);
                r.fingerprint.insert(r.fingerprint.end(), s.begin(), s.end());
            }
        } else {
            const ucfp_hash17 &h = algo_mask == UCFP_ALGO_PHASH ? out[i].phash : algo_mask == UCFP_ALGO_DHASH ? out[i].dhash : out[i].ahash;
            r.algorithm = algo_mask == UCFP_ALGO_PHASH ? ALGORITHM_PHASH : algo_mask == UCFP_ALGO_DHASH ? ALGORITHM_DHASH : ALGORITHM_AHASH;
            r.fingerprint = pack_single(imgs[i].exact32, h);
        }
    }
    return recs;
}
}  // namespace image

// IndexBackend for the scan path.  Storage stays with the host's redb backend; this mirrors vectors into HBM.
class GpuIndexBackend {
   public:
    explicit GpuIndexBackend(const Gpu &gpu, uint64_t capacity_per_tenant = 1u << 20) : gpu_(gpu), cap_(capacity_per_tenant) {}
    ~GpuIndexBackend() {
        for (auto &kv : vec_) ucfp_corpus_destroy(kv.second);
        for (auto &kv : ham_) ucfp_corpus_destroy(kv.second);
        for (auto &kv : mh_) ucfp_corpus_destroy(kv.second);
    }

    void upsert(const std::vector<Record> &batch) {  // src/index/mod.rs:20
        for (const Record &r : batch) {
            if (!r.embedding || r.embedding->empty()) continue;
            ucfp_corpus *&c = vec_[{r.tenant_id, r.embedding->size()}];
            if (!c) detail::check(ucfp_corpus_create(gpu_.raw(), UCFP_KIND_COSINE, (uint32_t)r.embedding->size(), cap_, &c), "Index");
            detail::check(ucfp_corpus_append(c, &r.record_id, r.embedding->data(), 1), "Index");
        }
    }

    // src/index/mod.rs:29-35 / embedded/mod.rs:268-360
    std::vector<Hit> knn(uint32_t tenant_id, const std::vector<float> &query, size_t k) const {
        std::vector<Hit> hits;
        if (query.empty() || k == 0) return hits;
        auto it = vec_.find({tenant_id, query.size()});
        if (it == vec_.end()) return hits;
        std::vector<uint64_t> ids(k);
        std::vector<float> scores(k);
        detail::check(ucfp_scan_cosine(it->second, query.data(), 1, k, ids.data(), scores.data()), "Index");
        for (size_t i = 0; i < k && ids[i] != UCFP_ID_NONE; ++i) hits.push_back(Hit{tenant_id, ids[i], scores[i], HitSource::Vector, {}, {}, {}, {}});
        return hits;
    }

    // ---- new: HashIndex (SURVEY 8b "new API the reference lacks") ----
    // 64-bit global hash of a stored image fingerprint (offset 32 of the 168-byte blob; PHash of a 536-byte bundle = offset 232)
    void upsert_hash(uint32_t tenant_id, const std::string &algorithm, uint64_t record_id, uint64_t code) {
        ucfp_corpus *&c = ham_[{tenant_id, algorithm}];
        if (!c) detail::check(ucfp_corpus_create(gpu_.raw(), UCFP_KIND_HAMMING64, 0, cap_, &c), "Index");
        detail::check(ucfp_corpus_append(c, &record_id, &code, 1), "Index");
    }
    void upsert_signature(uint32_t tenant_id, uint64_t record_id, const std::array<uint64_t, 128> &slots) {
        ucfp_corpus *&c = mh_[tenant_id];
        if (!c) detail::check(ucfp_corpus_create(gpu_.raw(), UCFP_KIND_MINHASH128, 0, cap_, &c), "Index");
        detail::check(ucfp_corpus_append(c, &record_id, slots.data(), 1), "Index");
    }
    // Hit.score = 1 - dist / 64
    std::vector<Hit> hamming_knn(uint32_t tenant_id, const std::string &algorithm, uint64_t code, size_t k) const {
        std::vector<Hit> hits;
        auto it = ham_.find({tenant_id, algorithm});
        if (it == ham_.end() || k == 0) return hits;
        std::vector<uint64_t> ids(k);
        std::vector<uint32_t> dist(k);
        detail::check(ucfp_scan_hamming(it->second, &code, 1, k, ids.data(), dist.data()), "Index");
        for (size_t i = 0; i < k && ids[i] != UCFP_ID_NONE; ++i)
            hits.push_back(Hit{tenant_id, ids[i], 1.0f - (float)dist[i] / 64.0f, HitSource::Vector, {}, {}, {}, {}});
        return hits;
    }
    // Hit.score = matches / 128
    std::vector<Hit> jaccard_knn(uint32_t tenant_id, const std::array<uint64_t, 128> &slots, size_t k) const {
        std::vector<Hit> hits;
        auto it = mh_.find(tenant_id);
        if (it == mh_.end() || k == 0) return hits;
        std::vector<uint64_t> ids(k);
        std::vector<uint32_t> matches(k);
        detail::check(ucfp_scan_jaccard(it->second, slots.data(), 1, k, ids.data(), matches.data()), "Index");
        for (size_t i = 0; i < k && ids[i] != UCFP_ID_NONE; ++i)
            hits.push_back(Hit{tenant_id, ids[i], (float)matches[i] / 128.0f, HitSource::Vector, {}, {}, {}, {}});
        return hits;
    }

   private:
    const Gpu &gpu_;
    uint64_t cap_;
    std::map<std::pair<uint32_t, size_t>, ucfp_corpus *> vec_;
    std::map<std::pair<uint32_t, std::string>, ucfp_corpus *> ham_;
    std::map<uint32_t, ucfp_corpus *> mh_;
};

class Matcher {  // src/matcher/mod.rs:140-207, vector arm; BM25 / hybrid stay on the host backend
   public:
    explicit Matcher(const GpuIndexBackend &index) : index_(index) {}
    std::vector<Hit> search(const Query &q) const {
        std::vector<Hit> fused;
        if (q.vector && q.terms.empty()) fused = index_.knn(q.tenant_id, *q.vector, q.k);
        else if (q.vector || !q.terms.empty()) throw Error("Unsupported", "bm25 / hybrid arms run on the host backend");
        else if (q.hash) fused = index_.hamming_knn(q.tenant_id, q.hash_algorithm, *q.hash, q.k);      // new arms, absent upstream
        else if (q.signature) fused = index_.jaccard_knn(q.tenant_id, *q.signature, q.k);
        if (fused.size() > q.k) fused.resize(q.k);
        return fused;
    }

   private:
    const GpuIndexBackend &index_;
};

}  // namespace ucfp
