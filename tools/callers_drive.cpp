// callers_drive.cpp — native load generator for the micro-batching front: N threads, each submitting ONE
// query per call through oc_batcher_search (the reference's one-search-per-task shape), so the number is
// not limited by the Python GIL.  Built by tools/bench_callers.py:
//   g++ -O2 -std=c++17 -shared -fPIC -pthread -Iinclude tools/callers_drive.cpp -o gpurun_out/libcallers_drive.so \
//       -Loramacore_b200 -loramacore_b200 -Wl,-rpath,$PWD/oramacore_b200
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "oramacore_b200.h"

extern "C" {
// Queries are given in CSR form (as oc_search_params would hold them for the whole set); every thread takes
// queries i = t, t + n_threads, ... for `rounds` passes.  Outputs: hits of the LAST pass ([Q][limit]), elapsed
// seconds of the timed region.  Returns the first non-zero oc status, or 0.
int callers_drive(oc_batcher *b, int mode, uint32_t n_threads, uint32_t rounds, uint32_t Q, uint32_t dim, uint32_t limit,
                  float similarity, float bm25_k, float bm25_b, const float *q_vecs, const uint32_t *q_token_offsets,
                  const uint32_t *token_term_offsets, const uint32_t *term_field, const uint32_t *term_id,
                  const float *term_weight, uint64_t *out_docs, float *out_scores, uint32_t *out_n, uint64_t *out_count,
                  double *elapsed_s) {
    std::atomic<int> rc{0};
    std::atomic<uint32_t> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; t++)
        th.emplace_back([&, t] {
            ready++;
            while (!go.load()) std::this_thread::yield();
            for (uint32_t r = 0; r < rounds && rc.load() == 0; r++)
                for (uint32_t i = t; i < Q; i += n_threads) {
                    oc_search_params p;
                    memset(&p, 0, sizeof(p));
                    p.mode = mode; p.n_queries = 1; p.limit = limit; p.offset = 0; p.similarity = similarity;
                    p.threshold = -1.0f; p.bm25_k = bm25_k; p.bm25_b = bm25_b;
                    if (mode != OC_MODE_FULLTEXT) p.q_vecs = q_vecs + size_t(i) * dim;
                    if (mode != OC_MODE_VECTOR) {   // the batcher honours non-zero-based offsets: point into the global CSR
                        p.q_token_offsets = q_token_offsets + i;
                        p.token_term_offsets = token_term_offsets;
                        p.term_field = term_field; p.term_id = term_id; p.term_weight = term_weight;
                    }
                    const int s = oc_batcher_search(b, &p, out_docs + size_t(i) * limit, out_scores + size_t(i) * limit,
                                                    out_n + i, out_count + i);
                    if (s != 0) { int z = 0; rc.compare_exchange_strong(z, s); break; }
                }
        });
    while (ready.load() < n_threads) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true);
    for (auto &x : th) x.join();
    *elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc.load();
}
}
