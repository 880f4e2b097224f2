/* driver.c — a plain C99 consumer of include/oramacore_b200.h (nothing else from this repo): loads the
 * committed case tests/golden/c_driver_case.bin (inputs + the oracle's answer), builds an embedding store and
 * a string store through the C ABI, runs ONE hybrid oc_search and compares.  This is the boundary a Rust /
 * Go / C host binds; the ctypes tests exercise the same entry points from Python.
 *   gcc -std=c99 -Wall -Wextra -pedantic -I include tests/c_driver/driver.c -o driver -L oramacore_b200 \
 *       -l:liboramacore_b200.so -Wl,-rpath,$PWD/oramacore_b200 && ./driver tests/golden/c_driver_case.bin */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oramacore_b200.h"

#define CHECK(x)                                                                       \
    do {                                                                               \
        int rc_ = (x);                                                                 \
        if (rc_ != OC_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, oc_last_error()); return 2; } \
    } while (0)

static void *take(const unsigned char **p, size_t bytes) {
    void *m = malloc(bytes ? bytes : 1);
    memcpy(m, *p, bytes);
    *p += bytes;
    return m;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: driver case.bin\n"); return 64; }
    FILE *fh = fopen(argv[1], "rb");
    if (!fh) { perror(argv[1]); return 66; }
    fseek(fh, 0, SEEK_END);
    long sz = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    unsigned char *blob = (unsigned char *)malloc((size_t)sz);
    if (fread(blob, 1, (size_t)sz, fh) != (size_t)sz) { fprintf(stderr, "short read\n"); return 66; }
    fclose(fh);
    const unsigned char *p = blob;
    uint32_t magic, n, dim, vocab, B, limit, n_tok, n_ent;
    uint64_t n_post;
    float avg_len;
    memcpy(&magic, p, 4); p += 4; memcpy(&n, p, 4); p += 4; memcpy(&dim, p, 4); p += 4; memcpy(&vocab, p, 4); p += 4;
    memcpy(&B, p, 4); p += 4; memcpy(&limit, p, 4); p += 4; memcpy(&n_post, p, 8); p += 8;
    memcpy(&n_tok, p, 4); p += 4; memcpy(&n_ent, p, 4); p += 4; memcpy(&avg_len, p, 4); p += 4;
    if (magic != 0x0C0DE001u) { fprintf(stderr, "bad magic\n"); return 65; }
    float *rows = (float *)take(&p, (size_t)n * dim * 4);
    uint64_t *term_offsets = (uint64_t *)take(&p, ((size_t)vocab + 1) * 8);
    uint32_t *post_row = (uint32_t *)take(&p, n_post * 4);
    uint16_t *post_tf = (uint16_t *)take(&p, n_post * 2);
    uint16_t *post_len = (uint16_t *)take(&p, n_post * 2);
    float *qv = (float *)take(&p, (size_t)B * dim * 4);
    uint32_t *q_tok = (uint32_t *)take(&p, ((size_t)B + 1) * 4);
    uint32_t *tok_term = (uint32_t *)take(&p, ((size_t)n_tok + 1) * 4);
    uint32_t *t_field = (uint32_t *)take(&p, (size_t)n_ent * 4);
    uint32_t *t_id = (uint32_t *)take(&p, (size_t)n_ent * 4);
    float *t_w = (float *)take(&p, (size_t)n_ent * 4);
    uint64_t *e_docs = (uint64_t *)take(&p, (size_t)B * limit * 8);
    float *e_scores = (float *)take(&p, (size_t)B * limit * 4);
    uint32_t *e_n = (uint32_t *)take(&p, (size_t)B * 4);
    uint64_t *e_cnt = (uint64_t *)take(&p, (size_t)B * 8);
    if (p - blob != sz) { fprintf(stderr, "layout mismatch: %ld of %ld bytes\n", (long)(p - blob), sz); return 65; }

    size_t abi[4];
    oc_abi_sizes(abi);
    if (abi[0] != sizeof(oc_search_params) || abi[1] != sizeof(oc_timing)) { fprintf(stderr, "ABI struct sizes differ\n"); return 3; }

    oc_ctx *ctx = NULL; oc_emb *emb = NULL; oc_str *str = NULL;
    CHECK(oc_init(0, &ctx));
    CHECK(oc_emb_create(ctx, dim, OC_DTYPE_F32, 0, &emb));
    uint64_t *ids = (uint64_t *)malloc((size_t)n * 8);
    for (uint32_t i = 0; i < n; i++) ids[i] = i;
    CHECK(oc_emb_insert(emb, ids, rows, n));
    CHECK(oc_str_create(ctx, 1, &str));
    CHECK(oc_str_set_rows(str, n, NULL, n));
    CHECK(oc_str_load_field(str, 0, avg_len, vocab, term_offsets, post_row, post_tf, post_len, NULL));

    oc_search_params sp;
    memset(&sp, 0, sizeof(sp));
    sp.mode = OC_MODE_HYBRID; sp.n_queries = B; sp.limit = limit; sp.offset = 0;
    sp.similarity = 0.0f; sp.threshold = -1.0f; sp.bm25_k = 1.2f; sp.bm25_b = 0.75f;
    sp.q_vecs = qv; sp.q_token_offsets = q_tok; sp.token_term_offsets = tok_term;
    sp.term_field = t_field; sp.term_id = t_id; sp.term_weight = t_w;
    uint64_t *docs = (uint64_t *)calloc((size_t)B * limit, 8), *cnt = (uint64_t *)calloc(B, 8);
    float *scores = (float *)calloc((size_t)B * limit, 4);
    uint32_t *nn = (uint32_t *)calloc(B, 4);
    CHECK(oc_search(ctx, emb, str, &sp, docs, scores, nn, cnt));

    int bad = 0;
    for (uint32_t q = 0; q < B; q++) {
        if (nn[q] != e_n[q] || cnt[q] != e_cnt[q]) { fprintf(stderr, "q%u: n %u/%u count %llu/%llu\n", q, nn[q], e_n[q], (unsigned long long)cnt[q], (unsigned long long)e_cnt[q]); bad++; continue; }
        for (uint32_t i = 0; i < nn[q]; i++) {
            const size_t k = (size_t)q * limit + i;
            if (docs[k] != e_docs[k] || fabsf(scores[k] - e_scores[k]) > 1e-5f) {
                fprintf(stderr, "q%u #%u: doc %llu/%llu score %.8f/%.8f\n", q, i, (unsigned long long)docs[k], (unsigned long long)e_docs[k], scores[k], e_scores[k]);
                bad++;
            }
        }
    }
    oc_timing t;
    CHECK(oc_last_timing(ctx, &t));
    printf("c_driver: %u queries, %u kernel launches, device %.3f ms, mismatches %d\n", B, t.kernel_launches, t.device_ms, bad);
    oc_str_destroy(str); oc_emb_destroy(emb); oc_shutdown(ctx);
    return bad ? 1 : 0;
}
