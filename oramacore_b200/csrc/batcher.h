// batcher.h — micro-batching front for concurrent single-query callers (host only, no CUDA).
//
// The reference runs ONE search per tokio task, many at a time (SURVEY.md §8b "Threading");
// the GPU path earns its throughput on batches.  SURVEY §8b allows "a batching queue" behind the
// boundary: threads submit one query each, the first submitter of a group becomes its leader,
// waits up to max_wait_us (or until max_batch queries are in), merges the group's query
// descriptors into ONE oc_search_params, runs it through `Exec` (oc_search in the library, a fake
// in tests/batcher_test.cpp) and scatters the per-query results back to the waiting callers.
// Only queries that can share a batch are coalesced: same (mode, limit, offset, similarity,
// threshold, bm25_k, bm25_b), no filter, no OMC, not sharded; anything else runs directly.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/oramacore_b200.h"

namespace ocb {

struct BatchKey {
    int mode; uint32_t limit, offset; float similarity, threshold, k, b; uint32_t vector_limit;
    bool operator==(const BatchKey &o) const {
        return mode == o.mode && limit == o.limit && offset == o.offset && vector_limit == o.vector_limit && memcmp(&similarity, &o.similarity, 4) == 0 &&
               memcmp(&threshold, &o.threshold, 4) == 0 && memcmp(&k, &o.k, 4) == 0 && memcmp(&b, &o.b, 4) == 0;
    }
};
inline BatchKey key_of(const oc_search_params *p) {
    return BatchKey{p->mode, p->limit, p->offset, p->similarity, p->threshold, p->bm25_k, p->bm25_b, p->vector_limit};
}
// has_emb / has_str: the stores the batcher was created with.  A call that oc_search would reject
// (unknown mode, missing store, NULL query arrays) is NOT batchable: it goes straight to the
// executor so the caller gets oc_search's normal error instead of a merge that dereferences NULL.
inline bool batchable(const oc_search_params *p, bool has_emb = true, bool has_str = true) {
    if (p->n_queries != 1 || p->filter_bits || p->filter || p->n_omc != 0 || p->sharded) return false;
    if (p->mode != OC_MODE_FULLTEXT && p->mode != OC_MODE_VECTOR && p->mode != OC_MODE_HYBRID) return false;
    const bool need_v = p->mode != OC_MODE_FULLTEXT, need_ft = p->mode != OC_MODE_VECTOR;
    if (need_v && (!has_emb || !p->q_vecs)) return false;
    if (need_ft) {
        if (!has_str || !p->q_token_offsets) return false;
        const uint32_t t0 = p->q_token_offsets[0], t1 = p->q_token_offsets[1];
        if (t1 < t0) return false;
        if (t1 > t0 && !p->token_term_offsets) return false;
        if (t1 > t0 && p->token_term_offsets[t1] > p->token_term_offsets[t0] && (!p->term_field || !p->term_id)) return false;
    }
    return true;
}

struct BatchReq {
    const oc_search_params *p;
    uint64_t *docs; float *scores; uint32_t *n; uint64_t *count;
    int rc = 0;
    bool done = false;
};

// Merged descriptors of one batch (owns the concatenated arrays the merged params point into).
struct MergedBatch {
    oc_search_params p{};
    std::vector<float> q_vecs, term_weight;
    std::vector<uint32_t> q_token_offsets, token_term_offsets, term_field, term_id;
    std::vector<uint64_t> docs, count;
    std::vector<float> scores;
    std::vector<uint32_t> n;

    void build(const std::vector<BatchReq *> &reqs, uint32_t dim) {
        const oc_search_params *f = reqs[0]->p;
        const uint32_t B = (uint32_t)reqs.size();
        p = *f;
        p.n_queries = B;
        const bool has_v = f->mode != OC_MODE_FULLTEXT, has_ft = f->mode != OC_MODE_VECTOR;
        if (has_v) {
            q_vecs.resize(size_t(B) * dim);
            for (uint32_t i = 0; i < B; i++) memcpy(q_vecs.data() + size_t(i) * dim, reqs[i]->p->q_vecs, size_t(dim) * 4);
            p.q_vecs = q_vecs.data();
        }
        if (has_ft) {
            q_token_offsets.assign(1, 0u);
            token_term_offsets.assign(1, 0u);
            for (uint32_t i = 0; i < B; i++) {
                const oc_search_params *r = reqs[i]->p;
                const uint32_t t0 = r->q_token_offsets[0], t1 = r->q_token_offsets[1];
                for (uint32_t t = t0; t < t1; t++) {
                    const uint32_t e0 = r->token_term_offsets[t], e1 = r->token_term_offsets[t + 1];
                    for (uint32_t e = e0; e < e1; e++) {
                        term_field.push_back(r->term_field[e]);
                        term_id.push_back(r->term_id[e]);
                        term_weight.push_back(r->term_weight ? r->term_weight[e] : 1.0f);
                    }
                    token_term_offsets.push_back((uint32_t)term_id.size());
                }
                q_token_offsets.push_back((uint32_t)token_term_offsets.size() - 1);
            }
            p.q_token_offsets = q_token_offsets.data();
            p.token_term_offsets = token_term_offsets.data();
            // empty vectors still need non-NULL pointers for the ABI's argument checks
            static const uint32_t zero_u = 0; static const float one_f = 1.0f;
            p.term_field = term_field.empty() ? &zero_u : term_field.data();
            p.term_id = term_id.empty() ? &zero_u : term_id.data();
            p.term_weight = term_weight.empty() ? &one_f : term_weight.data();
        }
        docs.assign(size_t(B) * f->limit, 0); scores.assign(size_t(B) * f->limit, 0.f);
        n.assign(B, 0); count.assign(B, 0);
    }
    void scatter(const std::vector<BatchReq *> &reqs, int rc) const {
        const uint32_t L = p.limit;
        for (size_t i = 0; i < reqs.size(); i++) {
            BatchReq *r = reqs[i];
            r->rc = rc;
            if (rc != 0) continue;
            memcpy(r->docs, docs.data() + i * L, size_t(L) * 8);
            memcpy(r->scores, scores.data() + i * L, size_t(L) * 4);
            *r->n = n[i];
            *r->count = count[i];
        }
    }
};

template <class Exec>   // int Exec(const oc_search_params*, uint64_t* docs, float* scores, uint32_t* n, uint64_t* count)
class Batcher {
public:
    Batcher(Exec exec, uint32_t dim, uint32_t max_batch, uint32_t max_wait_us, bool has_emb = true, bool has_str = true)
        : exec_(exec), dim_(dim), max_batch_(max_batch ? max_batch : 1), max_wait_us_(max_wait_us), has_emb_(has_emb), has_str_(has_str) {}

    int submit(const oc_search_params *p, uint64_t *docs, float *scores, uint32_t *n, uint64_t *count) {
        if (!batchable(p, has_emb_, has_str_) || max_batch_ == 1) {
            direct_++;
            return exec_(p, docs, scores, n, count);
        }
        BatchReq r{p, docs, scores, n, count};
        std::unique_lock<std::mutex> lk(mu_);
        // one group collects at a time: wait while it is full or holds a different parameter tuple
        const BatchKey k = key_of(p);
        cv_slot_.wait(lk, [&] { return pending_.empty() || (pending_key_ == k && pending_.size() < max_batch_); });
        if (pending_.empty()) pending_key_ = k;
        pending_.push_back(&r);
        if (!leader_active_) {
            leader_active_ = true;
            const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(max_wait_us_);
            while (pending_.size() < max_batch_)
                if (cv_leader_.wait_until(lk, deadline) == std::cv_status::timeout) break;
            std::vector<BatchReq *> batch;
            batch.swap(pending_);
            leader_active_ = false;          // the next arrival leads the next group while this one runs
            cv_slot_.notify_all();
            lk.unlock();
            MergedBatch m;
            m.build(batch, dim_);
            const int rc = exec_(&m.p, m.docs.data(), m.scores.data(), m.n.data(), m.count.data());
            m.scatter(batch, rc);
            lk.lock();
            batches_++; queries_ += batch.size();
            for (BatchReq *b : batch) b->done = true;
            cv_done_.notify_all();
        } else {
            if (pending_.size() >= max_batch_) cv_leader_.notify_one();
            cv_done_.wait(lk, [&] { return r.done; });
        }
        return r.rc;
    }
    void stats(uint64_t *queries, uint64_t *batches, uint64_t *direct) {
        std::lock_guard<std::mutex> g(mu_);
        if (queries) *queries = queries_;
        if (batches) *batches = batches_;
        if (direct) *direct = direct_.load();
    }

private:
    Exec exec_;
    uint32_t dim_, max_batch_, max_wait_us_;
    bool has_emb_, has_str_;
    std::mutex mu_;
    std::condition_variable cv_slot_, cv_leader_, cv_done_;
    std::vector<BatchReq *> pending_;
    BatchKey pending_key_{};
    bool leader_active_ = false;
    uint64_t queries_ = 0, batches_ = 0;
    std::atomic<uint64_t> direct_{0};
};

}  // namespace ocb
