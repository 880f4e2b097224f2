// oramacore_b200.hpp — header-only C++ mirror of the reference's read-side scoring interface
// over the C ABI (include/oramacore_b200.h).  The reference is compiled Rust; with no Rust
// toolchain in the build image the host side above the C ABI is C++ — same names, argument
// meaning and error behaviour (errors surface as exceptions carrying oc_last_error(), the
// analogue of anyhow::Error -> ReadError::Generic, read/mod.rs:137-138).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "oramacore_b200.h"

namespace oramacore_b200 {

struct Error : std::runtime_error {
    int code;
    Error(int c) : std::runtime_error(std::string("oramacore_b200 error ") + std::to_string(c) + ": " + oc_last_error()), code(c) {}
};
inline void check(int rc) { if (rc != OC_OK) throw Error(rc); }

using DocumentId = uint64_t;   // types.rs:111-112

class Context {                // one GPU (one process per GPU)
public:
    explicit Context(int device = 0) { check(oc_init(device, &h_)); }
    ~Context() { oc_shutdown(h_); }
    Context(const Context &) = delete;
    oc_ctx *get() const { return h_; }
private:
    oc_ctx *h_ = nullptr;
};

// committed_field/vector.rs:10-15
struct VectorSearchParams {
    const float *target;
    float similarity;
    size_t limit;
    const uint64_t *filtered_doc_ids = nullptr;   // bitmap over DocumentId (FilterResult::contains)
    uint64_t filter_nbits = 0;
};

// read/index/embedding_field.rs:29-34
class EmbeddingFieldStorage {
public:
    EmbeddingFieldStorage(Context &ctx, uint32_t dimensions, bool is_e5) { check(oc_emb_create(ctx.get(), dimensions, OC_DTYPE_F32, is_e5, &h_)); dim_ = dimensions; }
    ~EmbeddingFieldStorage() { oc_emb_destroy(h_); }
    // insert(DocumentId, Vec<Vec<f32>>) :232-237
    void insert(DocumentId doc, const std::vector<std::vector<float>> &vectors) {
        std::vector<float> flat; flat.reserve(vectors.size() * dim_);
        for (auto &v : vectors) flat.insert(flat.end(), v.begin(), v.end());
        std::vector<uint64_t> ids(vectors.size(), doc);
        check(oc_emb_insert(h_, ids.data(), flat.data(), ids.size()));
    }
    void remove(DocumentId doc) { check(oc_emb_delete(h_, &doc, 1)); }   // delete :240-242
    // search(&VectorSearchParams, &mut HashMap<DocumentId,f32>) :250-278 — output[doc] += score
    void search(const VectorSearchParams &p, std::unordered_map<DocumentId, float> &output) {
        std::vector<uint64_t> docs(p.limit); std::vector<float> scores(p.limit); uint32_t n = 0;
        check(oc_emb_search(h_, p.target, 1, (uint32_t)p.limit, p.similarity, p.filtered_doc_ids, p.filter_nbits,
                            docs.data(), scores.data(), &n));
        for (uint32_t i = 0; i < n; i++) output[docs[i]] += scores[i];
    }
    oc_emb *get() const { return h_; }
private:
    oc_emb *h_ = nullptr; uint32_t dim_ = 0;
};

struct TokenScore { DocumentId document_id; float score; };   // types.rs:362-366

// TokenScoreContext::execute + OMC + count + top-N for a batch (token_score.rs:460-509,
// search.rs:39-48, 482-498, sort.rs:260-279)
struct SearchOutput { std::vector<std::vector<TokenScore>> hits; std::vector<uint64_t> count; };
inline SearchOutput search(Context &ctx, oc_emb *emb, oc_str *str, const oc_search_params &p) {
    const uint32_t B = p.n_queries;
    std::vector<uint64_t> docs(size_t(B) * p.limit), cnt(B); std::vector<float> sc(size_t(B) * p.limit); std::vector<uint32_t> n(B);
    check(oc_search(ctx.get(), emb, str, &p, docs.data(), sc.data(), n.data(), cnt.data()));
    SearchOutput o; o.hits.resize(B); o.count = cnt;
    for (uint32_t q = 0; q < B; q++)
        for (uint32_t i = 0; i < n[q]; i++) o.hits[q].push_back({docs[size_t(q) * p.limit + i], sc[size_t(q) * p.limit + i]});
    return o;
}

// Micro-batching front for request tasks that call one search each (oc_batcher_*): thread-safe, blocks the
// caller for its batch's latency.  p.n_queries must be 1.
class SearchBatcher {
public:
    SearchBatcher(Context &ctx, oc_emb *emb, oc_str *str, uint32_t max_batch = 256, uint32_t max_wait_us = 200) {
        check(oc_batcher_create(ctx.get(), emb, str, max_batch, max_wait_us, &h_));
    }
    ~SearchBatcher() { oc_batcher_destroy(h_); }
    SearchBatcher(const SearchBatcher &) = delete;
    SearchBatcher &operator=(const SearchBatcher &) = delete;
    std::vector<TokenScore> search(const oc_search_params &p, uint64_t *count = nullptr) {
        std::vector<uint64_t> docs(p.limit); std::vector<float> sc(p.limit);
        uint32_t n = 0; uint64_t cnt = 0;
        check(oc_batcher_search(h_, &p, docs.data(), sc.data(), &n, &cnt));
        if (count) *count = cnt;
        std::vector<TokenScore> out;
        for (uint32_t i = 0; i < n; i++) out.push_back({docs[i], sc[i]});
        return out;
    }
private:
    oc_batcher *h_ = nullptr;
};

}  // namespace oramacore_b200
