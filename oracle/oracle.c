/*
 * oracle/oracle.c — CPU restatement of OramaCore's search hot path (see oracle.h header:
 * TEST INFRASTRUCTURE ONLY; never loaded by the product path).
 *
 * Compile with -ffp-contract=off: the reference is scalar Rust fp32 without FMA contraction.
 * Data-structure class follows the reference: hash-map score accumulation per query
 * (token_score.rs:257-300, bm25.rs:484-520) and a capped binary heap for top-N (sort.rs:260-279).
 */
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ scalar pieces */

/* bm25.rs:78-82  calculate_idf: ln_1p((N - df + 0.5) / (df + 0.5)) */
float orc_idf(float total_documents, uint64_t corpus_df) {
    float df = (float)corpus_df;
    float ratio = (total_documents - df + 0.5f) / (df + 0.5f);
    return log1pf(ratio);
}

/* bm25.rs:99-110  bm25f_normalized_tf: tf / (1 - b + b * (len / avglen)) */
float orc_normalized_tf(uint32_t tf, uint32_t field_len, float avg_field_len, float b) {
    float tff = (float)tf, len = (float)field_len;
    return tff / (1.0f - b + b * (len / avg_field_len));
}

/* bm25.rs:124-126  bm25f_score: idf * (k + 1) * S / (k + S) */
float orc_bm25f_score(float aggregated, float k, float idf) {
    return idf * (k + 1.0f) * aggregated / (k + aggregated);
}

/* bm25.rs:248-310  BM25Scorer::add (legacy one-call path) */
float orc_bm25_legacy_add(uint32_t tf, uint32_t field_len, float avg_len, float total_docs,
                          uint64_t df, float k, float weight, float b, float boost) {
    float ntf = orc_normalized_tf(tf, field_len, avg_len, b);
    float weighted = weight * ntf;
    float idf = orc_idf(total_docs, df);
    float s = orc_bm25f_score(weighted, k, idf);
    if (isnan(s)) return s;
    return s * boost;
}

/* python/embeddings.rs:71-92  Model::rescale_score */
float orc_rescale_score(float score, int is_e5) {
    if (!is_e5) return score;
    const float MIN = 0.7f, MAX = 1.0f, DELTA = MAX - MIN;
    float c = score;
    /* f32::clamp: NaN stays NaN */
    if (c < MIN) c = MIN;
    if (c > MAX) c = MAX;
    return (c - MIN) / DELTA;
}

static inline int f32_is_normal(float x) { return fpclassify(x) == FP_NORMAL; }

static inline int filter_contains(const uint64_t *bits, uint64_t nbits, uint64_t doc) {
    if (!bits) return 1;
    if (doc >= nbits) return 0;
    return (int)((bits[doc >> 6] >> (doc & 63)) & 1u);
}

/* ------------------------------------------------------------------ hash map u64 -> {f32,u32} */

typedef struct {
    uint64_t *key;   /* UINT64_MAX = empty */
    float *val;
    uint32_t *aux;
    size_t cap, n;   /* cap power of two */
} hmap;

static int hmap_init(hmap *m, size_t want) {
    size_t cap = 16;
    while (cap < want * 2) cap <<= 1;
    m->key = (uint64_t *)malloc(cap * sizeof(uint64_t));
    m->val = (float *)malloc(cap * sizeof(float));
    m->aux = (uint32_t *)malloc(cap * sizeof(uint32_t));
    if (!m->key || !m->val || !m->aux) return -1;
    memset(m->key, 0xff, cap * sizeof(uint64_t));
    m->cap = cap;
    m->n = 0;
    return 0;
}
static void hmap_free(hmap *m) {
    free(m->key); free(m->val); free(m->aux);
    m->key = NULL; m->val = NULL; m->aux = NULL; m->cap = m->n = 0;
}
static void hmap_clear(hmap *m) {
    if (m->n) memset(m->key, 0xff, m->cap * sizeof(uint64_t));
    m->n = 0;
}
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static int hmap_grow(hmap *m);
/* returns slot index; *fresh = 1 when newly inserted (val/aux zeroed) */
static inline size_t hmap_entry(hmap *m, uint64_t k, int *fresh) {
    if ((m->n + 1) * 4 > m->cap * 3) hmap_grow(m);
    size_t mask = m->cap - 1, i = (size_t)mix64(k) & mask;
    for (;;) {
        if (m->key[i] == k) { *fresh = 0; return i; }
        if (m->key[i] == UINT64_MAX) {
            m->key[i] = k; m->val[i] = 0.0f; m->aux[i] = 0; m->n++;
            *fresh = 1; return i;
        }
        i = (i + 1) & mask;
    }
}
static inline long hmap_find(const hmap *m, uint64_t k) {
    if (!m->cap) return -1;
    size_t mask = m->cap - 1, i = (size_t)mix64(k) & mask;
    for (;;) {
        if (m->key[i] == k) return (long)i;
        if (m->key[i] == UINT64_MAX) return -1;
        i = (i + 1) & mask;
    }
}
static int hmap_grow(hmap *m) {
    hmap o = *m;
    if (hmap_init(m, o.cap) != 0) return -1;   /* doubles */
    for (size_t i = 0; i < o.cap; i++)
        if (o.key[i] != UINT64_MAX) {
            int f; size_t j = hmap_entry(m, o.key[i], &f);
            m->val[j] = o.val[i]; m->aux[j] = o.aux[i];
        }
    hmap_free(&o);
    return 0;
}

/* ------------------------------------------------------------------ orc_map helpers */

void orc_map_free(orc_map *m) {
    if (!m) return;
    free(m->doc); free(m->score);
    m->doc = NULL; m->score = NULL; m->n = m->cap = 0;
}

typedef struct { uint64_t d; float s; } ds_pair;
static int cmp_doc(const void *a, const void *b) {
    uint64_t x = ((const ds_pair *)a)->d, y = ((const ds_pair *)b)->d;
    return x < y ? -1 : x > y;
}
static int hmap_to_sorted(const hmap *h, orc_map *out) {
    out->n = 0; out->cap = h->n;
    out->doc = (uint64_t *)malloc((h->n ? h->n : 1) * sizeof(uint64_t));
    out->score = (float *)malloc((h->n ? h->n : 1) * sizeof(float));
    ds_pair *tmp = (ds_pair *)malloc((h->n ? h->n : 1) * sizeof(ds_pair));
    if (!out->doc || !out->score || !tmp) { free(tmp); return -1; }
    size_t n = 0;
    for (size_t i = 0; i < h->cap; i++)
        if (h->key[i] != UINT64_MAX) { tmp[n].d = h->key[i]; tmp[n].s = h->val[i]; n++; }
    qsort(tmp, n, sizeof(ds_pair), cmp_doc);
    for (size_t i = 0; i < n; i++) { out->doc[i] = tmp[i].d; out->score[i] = tmp[i].s; }
    out->n = n;
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------ full text */

/* search_full_text (token_score.rs:186-303).  `scores` receives the final map
 * (get_scores, bm25.rs:416-428 / 522-524). */
static int fulltext_core(const orc_str_index *ix, const orc_text_query *q, const orc_text_params *p,
                         hmap *scores) {
    const int with_threshold = p->threshold >= 0.0f;
    uint32_t required = 0;
    if (with_threshold) {
        /* token_score.rs:211-218: perc = tokens.len() as f32 * threshold; floor as u32 */
        float perc = (float)q->n_tokens * p->threshold;
        required = (uint32_t)floorf(perc);
    }
    const float total_documents = (float)ix->document_count; /* token_score.rs:221 */
    hmap cur; /* current_term_contributions: doc -> S (val), aux unused; also = corpus_docs set */
    if (hmap_init(&cur, 1024) != 0) return -1;

    for (uint32_t ti = 0; ti < q->n_tokens; ti++) {
        hmap_clear(&cur); /* scorer.reset_term()/next_term() */
        uint64_t shard_df = 0; /* sharded corpus: df of a single-term token comes from the global table */
        if (q->token_term_offsets[ti + 1] - q->token_term_offsets[ti] == 1) {
            uint32_t e0 = q->token_term_offsets[ti];
            const orc_field *f0 = &ix->fields[q->term_field[e0]];
            if (f0->global_df && q->term_id[e0] < f0->n_terms) shard_df = f0->global_df[q->term_id[e0]];
        }
        for (uint32_t e = q->token_term_offsets[ti]; e < q->token_term_offsets[ti + 1]; e++) {
            const orc_field *f = &ix->fields[q->term_field[e]];
            uint32_t tid = q->term_id[e];
            if (tid >= f->n_terms) continue;
            float w = q->term_weight[e];
            for (uint64_t pi = f->term_offsets[tid]; pi < f->term_offsets[tid + 1]; pi++) {
                uint32_t row = f->post_row[pi];
                uint64_t doc = ix->row_doc_ids ? ix->row_doc_ids[row] : (uint64_t)row;
                if (!filter_contains(p->filter_bits, p->filter_nbits, doc)) continue;
                /* ntf "already includes boost + length normalization + exact_match_boost"
                 * (token_score.rs:180-185): ntf = w * tf / (1 - b + b*len/avglen), bm25.rs:99-110 */
                float ntf = w * orc_normalized_tf(f->post_tf[pi], f->post_len[pi], f->avg_field_len, p->b);
                int fresh; size_t s = hmap_entry(&cur, doc, &fresh); /* corpus_docs.insert(doc) */
                /* add_precomputed_field(doc, ntf, 1.0); S = sum(weight * ntf) in push order */
                cur.val[s] = cur.val[s] + 1.0f * ntf;
            }
        }
        uint64_t corpus_df = cur.n > 1 ? cur.n : 1; /* corpus_docs.len().max(1), token_score.rs:275 */
        if (shard_df) corpus_df = shard_df;
        float idf = orc_idf(total_documents, corpus_df);
        uint32_t bit = 1u << (ti & 31u); /* 1 << term_index (release-mode wrapping), token_score.rs:293 */
        for (size_t i = 0; i < cur.cap; i++) {
            if (cur.key[i] == UINT64_MAX) continue;
            float S = cur.val[i];
            if (!f32_is_normal(S)) continue;                 /* bm25.rs:387, 501 */
            float term_score = orc_bm25f_score(S, p->k, idf);
            if (isnan(term_score)) continue;                 /* bm25.rs:391, 505 */
            float final_score = term_score * 1.0f;           /* phrase_boost = 1.0 */
            int fresh; size_t s = hmap_entry(scores, cur.key[i], &fresh);
            scores->val[s] += final_score;
            scores->aux[s] |= bit;
        }
    }
    hmap_free(&cur);

    if (with_threshold) { /* bm25.rs:416-428: keep popcount(mask) >= threshold */
        hmap kept;
        if (hmap_init(&kept, scores->n) != 0) return -1;
        for (size_t i = 0; i < scores->cap; i++) {
            if (scores->key[i] == UINT64_MAX) continue;
            if ((uint32_t)__builtin_popcount(scores->aux[i]) >= required) {
                int fresh; size_t s = hmap_entry(&kept, scores->key[i], &fresh);
                kept.val[s] = scores->val[i]; kept.aux[s] = scores->aux[i];
            }
        }
        hmap_free(scores);
        *scores = kept;
    }
    return 0;
}

int orc_fulltext(const orc_str_index *ix, const orc_text_query *q, const orc_text_params *p,
                 orc_map *out) {
    hmap scores;
    if (hmap_init(&scores, 1024) != 0) return -1;
    int rc = fulltext_core(ix, q, p, &scores);
    if (rc == 0) rc = hmap_to_sorted(&scores, out);
    hmap_free(&scores);
    return rc;
}

/* ------------------------------------------------------------------ vectors */

/* fp32 dot product the way a SIMD CPU loop does it: 4 x 8 independent lane accumulators
 * (an AVX2 loop unrolled 4x), separate mul and add (no FMA), then a fixed reduction tree. */
typedef float v8f __attribute__((vector_size(32), aligned(4)));
static inline float dot8(const float *a, const float *b, uint32_t d) {
    v8f acc0 = {0}, acc1 = {0}, acc2 = {0}, acc3 = {0};
    uint32_t i = 0;
    for (; i + 32 <= d; i += 32) {
        acc0 = acc0 + *(const v8f *)(a + i) * *(const v8f *)(b + i);
        acc1 = acc1 + *(const v8f *)(a + i + 8) * *(const v8f *)(b + i + 8);
        acc2 = acc2 + *(const v8f *)(a + i + 16) * *(const v8f *)(b + i + 16);
        acc3 = acc3 + *(const v8f *)(a + i + 24) * *(const v8f *)(b + i + 24);
    }
    for (; i + 8 <= d; i += 8) acc0 = acc0 + *(const v8f *)(a + i) * *(const v8f *)(b + i);
    v8f acc = (acc0 + acc1) + (acc2 + acc3);
    float s = ((acc[0] + acc[4]) + (acc[1] + acc[5])) + ((acc[2] + acc[6]) + (acc[3] + acc[7]));
    for (; i < d; i++) s = s + a[i] * b[i];
    return s;
}

void orc_row_norms(const float *rows, uint64_t n_rows, uint32_t dim, float *out) {
    for (uint64_t r = 0; r < n_rows; r++) out[r] = sqrtf(dot8(rows + (size_t)r * dim, rows + (size_t)r * dim, dim));
}

typedef struct { float key; uint64_t tie; uint64_t payload; } hitem;
/* min-heap on (key asc, tie desc): root = the worst of the kept top set, where "better" is
 * larger key, then smaller tie. */
static inline int h_worse(const hitem *a, const hitem *b) {
    if (a->key != b->key) return a->key < b->key;
    return a->tie > b->tie;
}
static void h_sift_down(hitem *h, size_t n, size_t i) {
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && h_worse(&h[l], &h[m])) m = l;
        if (r < n && h_worse(&h[r], &h[m])) m = r;
        if (m == i) return;
        hitem t = h[i]; h[i] = h[m]; h[m] = t; i = m;
    }
}
static void h_sift_up(hitem *h, size_t i) {
    while (i) {
        size_t p = (i - 1) / 2;
        if (!h_worse(&h[i], &h[p])) return;
        hitem t = h[i]; h[i] = h[p]; h[p] = t; i = p;
    }
}
/* CappedHeap::insert (oramacore_lib, call sites sort.rs:203, 261-273) */
static inline void capped_insert(hitem *h, size_t *n, size_t cap, hitem it) {
    if (cap == 0) return;
    if (*n < cap) { h[*n] = it; h_sift_up(h, (*n)++); return; }
    if (h_worse(&h[0], &it)) { h[0] = it; h_sift_down(h, *n, 0); }
}
static int cmp_best_first(const void *a, const void *b) {
    const hitem *x = (const hitem *)a, *y = (const hitem *)b;
    if (h_worse(y, x)) return -1;
    if (h_worse(x, y)) return 1;
    return 0;
}

/* exact top-`limit` nearest by cosine distance = 1 - cos
 * (EmbeddingStorage::search contract, embedding_field.rs:246-266). Fills hits best-first. */
static size_t vector_topk(const orc_emb_store *st, const float *target, uint32_t limit,
                          const uint64_t *fbits, uint64_t fn, hitem *heap) {
    size_t hn = 0;
    const uint32_t d = st->dim;
    float qn = sqrtf(dot8(target, target, d));
    for (uint64_t r = 0; r < st->n_rows; r++) {
        if (st->deleted && st->deleted[r]) continue;
        uint64_t doc = st->row_doc_ids ? st->row_doc_ids[r] : r;
        if (!filter_contains(fbits, fn, doc)) continue;
        const float *x = st->rows + (size_t)r * d;
        float xn = st->row_norms ? st->row_norms[r] : sqrtf(dot8(x, x, d));
        float denom = xn * qn;
        float cosv = denom > 0.0f ? dot8(x, target, d) / denom : 0.0f;
        float distance = 1.0f - cosv;
        hitem it; it.key = -distance; it.tie = r; it.payload = doc;
        capped_insert(heap, &hn, limit, it);
    }
    qsort(heap, hn, sizeof(hitem), cmp_best_first);
    return hn;
}

/* EmbeddingFieldStorage::search (embedding_field.rs:250-278) into a hash map */
static int vector_core(const orc_emb_store *st, const float *target, uint32_t limit, float similarity,
                       const uint64_t *fbits, uint64_t fn, hmap *out) {
    hitem *heap = (hitem *)malloc((limit ? limit : 1) * sizeof(hitem));
    if (!heap) return -1;
    size_t hn = vector_topk(st, target, limit, fbits, fn, heap);
    for (size_t i = 0; i < hn; i++) {
        float distance = -heap[i].key;
        float sim = 1.0f - distance;                       /* :270 */
        float score = orc_rescale_score(sim, st->is_e5);   /* :271 */
        if (score >= similarity) {                         /* :272 */
            int fresh; size_t s = hmap_entry(out, heap[i].payload, &fresh);
            out->val[s] += score;                          /* :273-274 */
        }
    }
    free(heap);
    return 0;
}

int orc_vector(const orc_emb_store *st, const float *target, uint32_t limit, float similarity,
               const uint64_t *filter_bits, uint64_t filter_nbits, orc_map *out) {
    hmap h;
    if (hmap_init(&h, limit + 8) != 0) return -1;
    int rc = vector_core(st, target, limit, similarity, filter_bits, filter_nbits, &h);
    if (rc == 0) rc = hmap_to_sorted(&h, out);
    hmap_free(&h);
    return rc;
}

typedef struct { double key; uint64_t row; uint64_t doc; } ditem;
static int cmp_ditem(const void *a, const void *b) {
    const ditem *x = (const ditem *)a, *y = (const ditem *)b;
    if (x->key != y->key) return x->key > y->key ? -1 : 1;
    return x->row < y->row ? -1 : x->row > y->row;
}
int orc_vector_f64(const orc_emb_store *st, const float *target, uint32_t limit,
                   uint64_t *out_doc, double *out_cos) {
    const uint32_t d = st->dim;
    size_t cap = (size_t)limit * 2 + 64, n = 0;
    ditem *buf = (ditem *)malloc(cap * sizeof(ditem));
    if (!buf) return -1;
    double qn = 0;
    for (uint32_t i = 0; i < d; i++) qn += (double)target[i] * target[i];
    qn = sqrt(qn);
    double worst = -INFINITY;
    for (uint64_t r = 0; r < st->n_rows; r++) {
        if (st->deleted && st->deleted[r]) continue;
        const float *x = st->rows + (size_t)r * d;
        double dp = 0, xn = 0;
        for (uint32_t i = 0; i < d; i++) { dp += (double)x[i] * target[i]; xn += (double)x[i] * x[i]; }
        double c = (xn > 0 && qn > 0) ? dp / (sqrt(xn) * qn) : 0.0;
        if (n >= limit && c < worst) continue;
        buf[n].key = c; buf[n].row = r; buf[n].doc = st->row_doc_ids ? st->row_doc_ids[r] : r; n++;
        if (n == cap) {
            qsort(buf, n, sizeof(ditem), cmp_ditem);
            n = limit; worst = buf[n - 1].key;
        }
    }
    qsort(buf, n, sizeof(ditem), cmp_ditem);
    if (n > limit) n = limit;
    for (size_t i = 0; i < n; i++) { out_doc[i] = buf[i].doc; out_cos[i] = buf[i].key; }
    free(buf);
    return (int)n;
}

/* ------------------------------------------------------------------ fusion / omc / top-n */

/* normalize_and_combine (token_score.rs:393-422) on hash maps; result replaces `fulltext`. */
static void combine_core(const hmap *vector, hmap *fulltext) {
    float max = 0.0f, min = 0.0f;
    /* folds start at 0.0 (:398-401); f32::max/min ignore NaN */
    for (size_t i = 0; i < vector->cap; i++) if (vector->key[i] != UINT64_MAX) max = fmaxf(max, vector->val[i]);
    float m2 = 0.0f;
    for (size_t i = 0; i < fulltext->cap; i++) if (fulltext->key[i] != UINT64_MAX) m2 = fmaxf(m2, fulltext->val[i]);
    max = fmaxf(max, m2);
    for (size_t i = 0; i < vector->cap; i++) if (vector->key[i] != UINT64_MAX) min = fminf(min, vector->val[i]);
    m2 = 0.0f;
    for (size_t i = 0; i < fulltext->cap; i++) if (fulltext->key[i] != UINT64_MAX) m2 = fminf(m2, fulltext->val[i]);
    min = fminf(min, m2);
    for (size_t i = 0; i < fulltext->cap; i++)
        if (fulltext->key[i] != UINT64_MAX) fulltext->val[i] = (fulltext->val[i] - min) / (max - min);
    for (size_t i = 0; i < vector->cap; i++) {
        if (vector->key[i] == UINT64_MAX) continue;
        float v = (vector->val[i] - min) / (max - min);
        int fresh; size_t s = hmap_entry(fulltext, vector->key[i], &fresh);
        fulltext->val[s] += v; /* entry(k).or_default() += v */
    }
}

static int sorted_to_hmap(const orc_map *m, hmap *h) {
    if (hmap_init(h, m->n + 8) != 0) return -1;
    for (size_t i = 0; i < m->n; i++) { int f; size_t s = hmap_entry(h, m->doc[i], &f); h->val[s] = m->score[i]; }
    return 0;
}

int orc_hybrid_combine(const orc_map *vector, const orc_map *fulltext, orc_map *out) {
    hmap v, f;
    if (sorted_to_hmap(vector, &v) != 0 || sorted_to_hmap(fulltext, &f) != 0) return -1;
    combine_core(&v, &f);
    int rc = hmap_to_sorted(&f, out);
    hmap_free(&v); hmap_free(&f);
    return rc;
}

static long omc_find(const uint64_t *omc_doc, size_t n, uint64_t doc) {
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (omc_doc[mid] < doc) lo = mid + 1; else hi = mid; }
    return (lo < n && omc_doc[lo] == doc) ? (long)lo : -1;
}

/* apply_omc_multipliers (search.rs:39-48) */
void orc_apply_omc(orc_map *scores, const uint64_t *omc_doc, const float *omc_mult, size_t n_omc) {
    if (!n_omc) return;
    for (size_t i = 0; i < scores->n; i++) {
        long j = omc_find(omc_doc, n_omc, scores->doc[i]);
        if (j >= 0) scores->score[i] *= omc_mult[j];
    }
}
static void omc_core(hmap *scores, const uint64_t *omc_doc, const float *omc_mult, size_t n_omc) {
    if (!n_omc) return;
    for (size_t i = 0; i < scores->cap; i++) {
        if (scores->key[i] == UINT64_MAX) continue;
        long j = omc_find(omc_doc, n_omc, scores->key[i]);
        if (j >= 0) scores->val[i] *= omc_mult[j];
    }
}

/* top_n (sort.rs:260-279) over a hash map */
static size_t topn_core(const hmap *scores, size_t n, uint64_t *out_doc, float *out_score) {
    hitem *heap = (hitem *)malloc((n ? n : 1) * sizeof(hitem));
    size_t hn = 0;
    for (size_t i = 0; i < scores->cap; i++) {
        if (scores->key[i] == UINT64_MAX) continue;
        float v = scores->val[i];
        if (isnan(v)) continue;                /* NotNan::new(..) Err => continue */
        hitem it; it.key = v; it.tie = scores->key[i]; it.payload = scores->key[i];
        capped_insert(heap, &hn, n, it);
    }
    qsort(heap, hn, sizeof(hitem), cmp_best_first);
    for (size_t i = 0; i < hn; i++) { out_doc[i] = heap[i].payload; out_score[i] = heap[i].key; }
    free(heap);
    return hn;
}

size_t orc_top_n(const orc_map *scores, size_t n, uint64_t *out_doc, float *out_score) {
    hmap h;
    if (sorted_to_hmap(scores, &h) != 0) return 0;
    size_t r = topn_core(&h, n, out_doc, out_score);
    hmap_free(&h);
    return r;
}

/* ------------------------------------------------------------------ search() */

/* search_on_indexes restricted to the hot path (search.rs:283-501):
 * token scores -> OMC -> count -> top-(limit+offset) -> skip(offset).take(limit). */
int orc_search(const orc_str_index *ix, const orc_emb_store *st, const orc_search_req *r,
               uint64_t *out_doc, float *out_score, uint32_t *out_n, uint64_t *out_count) {
    hmap res;
    /* search.rs:121,297-298: map pre-sized to doc_count/3 */
    size_t est = ix ? (size_t)(ix->document_count / 3) : (st ? (size_t)(st->n_rows / 3) : 16);
    if (est > (1u << 22)) est = 1u << 22;
    if (hmap_init(&res, est + 16) != 0) return -1;
    int rc = 0;
    if (r->mode == 0) {
        rc = fulltext_core(ix, r->text, r->tp, &res);
    } else if (r->mode == 1) {
        rc = vector_core(st, r->q_vec, r->limit, r->similarity,
                         r->tp ? r->tp->filter_bits : NULL, r->tp ? r->tp->filter_nbits : 0, &res);
    } else {
        hmap vec;
        if (hmap_init(&vec, r->limit + 8) != 0) { hmap_free(&res); return -1; }
        rc = vector_core(st, r->q_vec, r->limit, r->similarity,
                         r->tp->filter_bits, r->tp->filter_nbits, &vec);   /* token_score.rs:368-374 */
        if (rc == 0) rc = fulltext_core(ix, r->text, r->tp, &res);         /* :375-383 */
        if (rc == 0) combine_core(&vec, &res);                             /* :386 */
        hmap_free(&vec);
    }
    if (rc != 0) { hmap_free(&res); return rc; }
    omc_core(&res, r->omc_doc, r->omc_mult, r->n_omc);   /* search.rs:342-343 */
    *out_count = res.n;                                  /* search.rs:482 */
    size_t want = (size_t)r->limit + r->offset;          /* sort.rs:25-26 */
    uint64_t *td = (uint64_t *)malloc((want ? want : 1) * sizeof(uint64_t));
    float *ts = (float *)malloc((want ? want : 1) * sizeof(float));
    size_t got = topn_core(&res, want, td, ts);
    uint32_t n = 0;
    for (size_t i = r->offset; i < got && n < r->limit; i++, n++) { /* search.rs:494-498 */
        out_doc[n] = td[i]; out_score[n] = ts[i];
    }
    *out_n = n;
    free(td); free(ts);
    hmap_free(&res);
    return 0;
}

typedef struct {
    const orc_str_index *ix; const orc_emb_store *st; const orc_search_req *reqs;
    uint32_t n_req; uint32_t *next; pthread_mutex_t *mu;
    uint64_t *out_doc; float *out_score; uint32_t *out_n; uint64_t *out_count; int rc;
    uint32_t stride;
} batch_arg;

static void *batch_worker(void *vp) {
    batch_arg *a = (batch_arg *)vp;
    for (;;) {
        pthread_mutex_lock(a->mu);
        uint32_t i = (*a->next)++;
        pthread_mutex_unlock(a->mu);
        if (i >= a->n_req) break;
        int rc = orc_search(a->ix, a->st, &a->reqs[i], a->out_doc + (size_t)i * a->stride,
                            a->out_score + (size_t)i * a->stride, &a->out_n[i], &a->out_count[i]);
        if (rc != 0) a->rc = rc;
    }
    return NULL;
}

int orc_search_batch(const orc_str_index *ix, const orc_emb_store *st, const orc_search_req *reqs,
                     uint32_t n_req, uint32_t n_threads, uint64_t *out_doc, float *out_score,
                     uint32_t *out_n, uint64_t *out_count) {
    if (n_threads == 0) n_threads = 1;
    uint32_t stride = 0;
    for (uint32_t i = 0; i < n_req; i++) if (reqs[i].limit > stride) stride = reqs[i].limit;
    uint32_t next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_t *th = (pthread_t *)malloc(n_threads * sizeof(pthread_t));
    batch_arg *args = (batch_arg *)malloc(n_threads * sizeof(batch_arg));
    for (uint32_t t = 0; t < n_threads; t++) {
        batch_arg a = {ix, st, reqs, n_req, &next, &mu, out_doc, out_score, out_n, out_count, 0, stride};
        args[t] = a;
        pthread_create(&th[t], NULL, batch_worker, &args[t]);
    }
    int rc = 0;
    for (uint32_t t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); if (args[t].rc) rc = args[t].rc; }
    free(th); free(args);
    return rc;
}
