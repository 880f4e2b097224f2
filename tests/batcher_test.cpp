// Host-logic test of the micro-batching queue (oramacore_b200/csrc/batcher.h) with a fake executor:
// 16 threads submit single queries with mixed parameter tuples; every caller must get exactly the
// result of ITS query (merge + scatter under concurrency), and queries must actually be coalesced.
// Built and run by tests/test_batcher_host.py (g++, no CUDA).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>

#include "../oramacore_b200/csrc/batcher.h"

static const uint32_t DIM = 8;

static long long signature(const oc_search_params *p, uint32_t i) {
    long long sig = 0;
    if (p->mode != OC_MODE_FULLTEXT) sig += (long long)llround(p->q_vecs[size_t(i) * DIM]) * 1000003LL;
    if (p->mode != OC_MODE_VECTOR) {
        for (uint32_t t = p->q_token_offsets[i]; t < p->q_token_offsets[i + 1]; t++)
            for (uint32_t e = p->token_term_offsets[t]; e < p->token_term_offsets[t + 1]; e++)
                sig += (long long)p->term_field[e] * 131 + (long long)p->term_id[e] * 7 +
                       (long long)llround((p->term_weight ? p->term_weight[e] : 1.0f) * 4) + 13LL * (t - p->q_token_offsets[i]);
    }
    return sig;
}
static uint64_t n_terms_of(const oc_search_params *p, uint32_t i) {
    if (p->mode == OC_MODE_VECTOR) return 0;
    const uint32_t t0 = p->q_token_offsets[i], t1 = p->q_token_offsets[i + 1];
    return (uint64_t)(p->token_term_offsets[t1] - p->token_term_offsets[t0]) + 1000ull * (t1 - t0);
}

struct FakeExec {
    std::atomic<int> *max_seen;
    int operator()(const oc_search_params *p, uint64_t *docs, float *scores, uint32_t *n, uint64_t *count) const {
        int prev = max_seen->load();
        while ((int)p->n_queries > prev && !max_seen->compare_exchange_weak(prev, (int)p->n_queries)) {}
        std::this_thread::sleep_for(std::chrono::microseconds(300));   // "device time": lets the next group fill up
        for (uint32_t i = 0; i < p->n_queries; i++) {
            const long long sig = signature(p, i);
            for (uint32_t j = 0; j < p->limit; j++) {
                docs[size_t(i) * p->limit + j] = (uint64_t)(sig * 1000 + j);
                scores[size_t(i) * p->limit + j] = (float)(sig % 1000) + 0.5f * j + p->similarity;
            }
            n[i] = p->limit;
            count[i] = n_terms_of(p, i);
        }
        return 0;
    }
};

int main() {
    std::atomic<int> max_seen{0};
    ocb::Batcher<FakeExec> b(FakeExec{&max_seen}, DIM, 32, 2000);
    std::atomic<int> bad{0};
    const int T = 16, Q = 250;
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            std::mt19937 rng(1234 + t);
            for (int it = 0; it < Q; it++) {
                oc_search_params p{};
                p.mode = (int)(rng() % 3);
                p.n_queries = 1;
                p.limit = (rng() % 2) ? 10 : 5;
                p.offset = 0;
                p.similarity = (rng() % 2) ? 0.0f : 0.7f;
                p.threshold = -1.0f; p.bm25_k = 1.2f; p.bm25_b = 0.75f;
                float qv[DIM];
                for (uint32_t d = 0; d < DIM; d++) qv[d] = (float)(rng() % 1000);
                p.q_vecs = qv;
                const uint32_t base = rng() % 3;                 // non-zero-based token offsets must be honoured
                const uint32_t ntok = rng() % 5;
                std::vector<uint32_t> qto = {base, base + ntok}, tto(base + ntok + 1, 0), tf, ti;
                std::vector<float> tw;
                for (uint32_t k = 0; k <= base; k++) tto[k] = 0;
                for (uint32_t k = 0; k < ntok; k++) {
                    const uint32_t nt = rng() % 4;                // 0 terms = unknown token
                    for (uint32_t e = 0; e < nt; e++) { tf.push_back(rng() % 3); ti.push_back(rng() % 5000); tw.push_back((float)(1 + rng() % 3)); }
                    tto[base + k + 1] = (uint32_t)ti.size();
                }
                static const uint32_t z = 0;
                p.q_token_offsets = qto.data(); p.token_term_offsets = tto.data();
                p.term_field = tf.empty() ? &z : tf.data(); p.term_id = ti.empty() ? &z : ti.data();
                const bool null_w = rng() % 4 == 0;
                if (null_w) for (auto &w : tw) w = 1.0f;
                p.term_weight = (null_w || tw.empty()) ? nullptr : tw.data();
                uint64_t dummy_filter = ~0ull;
                if (rng() % 10 == 0) { p.filter_bits = &dummy_filter; p.filter_nbits = 64; }   // not batchable: direct path
                std::vector<uint64_t> docs(p.limit); std::vector<float> sc(p.limit);
                uint32_t n = 0; uint64_t cnt = 0;
                const int rc = b.submit(&p, docs.data(), sc.data(), &n, &cnt);
                const long long sig = signature(&p, 0);
                bool ok = rc == 0 && n == p.limit && cnt == n_terms_of(&p, 0);
                for (uint32_t j = 0; ok && j < p.limit; j++)
                    ok = docs[j] == (uint64_t)(sig * 1000 + j) && sc[j] == (float)(sig % 1000) + 0.5f * j + p.similarity;
                if (!ok) bad++;
            }
        });
    for (auto &x : th) x.join();
    uint64_t q = 0, nb = 0, direct = 0;
    b.stats(&q, &nb, &direct);
    printf("queries=%llu batches=%llu direct=%llu max_batch_seen=%d bad=%d\n", (unsigned long long)q, (unsigned long long)nb,
           (unsigned long long)direct, max_seen.load(), bad.load());
    if (bad.load() != 0) return 1;
    if (q + direct != (uint64_t)T * Q) return 2;
    if (nb * 2 > q) return 3;          // coalescing must happen: on average >= 2 queries per batch
    if (max_seen.load() > 32) return 4;
    return 0;
}
