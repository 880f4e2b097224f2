/*
 * oracle/oracle.h — CPU restatement of OramaCore's search hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load this library.  The product
 * path (oramacore_b200/, include/oramacore_b200.h) never links, imports or calls it.
 *
 * Every function cites the reference file:line (relative to oramasearch/oramacore @ 666ab48)
 * whose arithmetic and operation order it restates.  Scores are computed in IEEE fp32 with
 * FMA contraction disabled, exactly like the reference's scalar Rust.
 *
 * Parity pin status:
 *   - BM25F arithmetic (idf / normalised tf / saturation / threshold mask): PINNED by the
 *     closed-form known-answer tests of src/collection_manager/bm25.rs:534-563, 912-983
 *     (tests/test_oracle_golden.py) and the ordering/count pins of
 *     src/tests/fulltext_search.rs:146-251, 478-600.
 *   - absolute cosine scores, hybrid fusion, top-k tie order: PARITY UNPINNED — the
 *     reference holds no golden vectors for them; the oracle restates the
 *     in-tree formulas and the documented contract `distance = 1 - cosine_similarity`.
 *   - the posting walk and the dense scan live in un-vendored crates
 *     (oramacore_fields 0.2.0, oramacore_lib 0.4.4; Cargo.lock:5313-5370); restated from
 *     their call sites and the in-tree statement of the ntf formula (bm25.rs:99-110).
 */
#ifndef ORAMACORE_ORACLE_H
#define ORAMACORE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar BM25F pieces (bm25.rs:78-82, 99-110, 124-126) ---- */
float orc_idf(float total_documents, uint64_t corpus_df);
float orc_normalized_tf(uint32_t tf, uint32_t field_len, float avg_field_len, float b);
float orc_bm25f_score(float aggregated, float k, float idf);
/* BM25Scorer::add legacy single-call path (bm25.rs:248-310) — used by the reference's
 * own test_bm25f_scorer_basic; returns the term score (NaN => skipped). */
float orc_bm25_legacy_add(uint32_t tf, uint32_t field_len, float avg_len, float total_docs,
                          uint64_t df, float k, float weight, float b, float boost);
/* Model::rescale_score (python/embeddings.rs:71-92) */
float orc_rescale_score(float similarity, int is_e5);

/* ---- string index (one StringFieldStorage per field; string_field.rs) ----
 * Postings in CSR by term; rows ascending inside a term; one posting per (term,row). */
typedef struct {
    float avg_field_len;          /* info().avg_field_length            */
    uint32_t n_terms;
    const uint64_t *term_offsets; /* n_terms + 1                        */
    const uint32_t *post_row;     /* row index into row_doc_ids         */
    const uint16_t *post_tf;      /* term frequency in this field       */
    const uint16_t *post_len;     /* field_length (u16, string_field.rs:162) */
    const uint32_t *global_df;    /* NULL, or per-term corpus df when this index is one shard of a
                                     document-sharded corpus (single-term tokens only)          */
} orc_field;

typedef struct {
    uint32_t n_fields;
    const orc_field *fields;
    uint64_t n_rows;
    const uint64_t *row_doc_ids;  /* NULL => doc_id == row; ascending   */
    uint64_t document_count;      /* N for idf (token_score.rs:221)     */
} orc_str_index;

/* One query, already resolved on the host: tokens -> expanded index terms.
 * (tokenise/stem/prefix/fuzzy expansion stays host-side, token_score.rs:196-209.) */
typedef struct {
    uint32_t n_tokens;
    const uint32_t *token_term_offsets; /* n_tokens + 1 */
    const uint32_t *term_field;         /* per expanded term: field index           */
    const uint32_t *term_id;            /* per expanded term: term id in that field */
    const float *term_weight;           /* boost * exact_match_boost (baked into ntf) */
} orc_text_query;

typedef struct {
    float b;            /* Bm25Params::default().b restated as 0.75 (bm25.rs:56-63) */
    float k;            /* 1.2 (token_score.rs:283,291) */
    float threshold;    /* <0 => plain scorer; else Threshold (token_score.rs:211-218) */
    const uint64_t *filter_bits; /* NULL => no filter; bit per doc_id */
    uint64_t filter_nbits;
} orc_text_params;

/* A score map: parallel arrays sorted by doc id (the reference's HashMap<DocumentId,f32>). */
typedef struct {
    uint64_t *doc;
    float *score;
    size_t n, cap;
} orc_map;
void orc_map_free(orc_map *m);

/* search_full_text (token_score.rs:186-303) + BM25Scorer (bm25.rs:325-524). */
int orc_fulltext(const orc_str_index *ix, const orc_text_query *q, const orc_text_params *p,
                 orc_map *out);

/* ---- embedding store (embedding_field.rs:232-278) ---- */
typedef struct {
    uint32_t dim;
    uint64_t n_rows;
    const float *rows;          /* n_rows x dim row-major fp32          */
    const uint64_t *row_doc_ids;/* NULL => doc_id == row                */
    const uint8_t *deleted;     /* NULL or n_rows flags                 */
    int is_e5;
    const float *row_norms;     /* NULL or precomputed |x| per row (orc_row_norms) */
} orc_emb_store;
/* |x| per row with the same fp32 blocked summation the scan uses. */
void orc_row_norms(const float *rows, uint64_t n_rows, uint32_t dim, float *out);

/* EmbeddingFieldStorage::search: exact top-`limit` by cosine distance, then
 * similarity = 1 - distance, rescale, keep >= similarity, output[doc] += score. */
int orc_vector(const orc_emb_store *st, const float *target, uint32_t limit, float similarity,
               const uint64_t *filter_bits, uint64_t filter_nbits, orc_map *out);
/* fp64 exact brute force (recall oracle): fills top-`limit` doc ids / cosine (double). */
int orc_vector_f64(const orc_emb_store *st, const float *target, uint32_t limit,
                   uint64_t *out_doc, double *out_cos);

/* normalize_and_combine (token_score.rs:393-422). Consumes nothing; writes `out`. */
int orc_hybrid_combine(const orc_map *vector, const orc_map *fulltext, orc_map *out);
/* apply_omc_multipliers (search.rs:39-48); omc sorted by doc id. */
void orc_apply_omc(orc_map *scores, const uint64_t *omc_doc, const float *omc_mult, size_t n_omc);
/* sort_token_scores/top_n (sort.rs:17-46, 260-279): NaN dropped, descending score,
 * ties broken by ascending doc id (reference tie order is unspecified). Returns n written. */
size_t orc_top_n(const orc_map *scores, size_t n, uint64_t *out_doc, float *out_score);

/* ---- whole search() for one query (search.rs:283-501 restricted to the hot path) ---- */
typedef struct {
    int mode;                 /* 0 fulltext/default, 1 vector, 2 hybrid */
    uint32_t limit, offset;
    float similarity;
    const float *q_vec;       /* dim floats or NULL */
    const orc_text_query *text;
    const orc_text_params *tp;
    const uint64_t *omc_doc; const float *omc_mult; size_t n_omc;
} orc_search_req;

int orc_search(const orc_str_index *ix, const orc_emb_store *st, const orc_search_req *r,
               uint64_t *out_doc, float *out_score, uint32_t *out_n, uint64_t *out_count);

/* Batch driver used as the timed CPU baseline: one query per thread over n_threads
 * (how the reference serves concurrent load; a single search is single-threaded,
 * SURVEY.md §3.1). Returns 0 on success. */
int orc_search_batch(const orc_str_index *ix, const orc_emb_store *st, const orc_search_req *reqs,
                     uint32_t n_req, uint32_t n_threads, uint64_t *out_doc, float *out_score,
                     uint32_t *out_n, uint64_t *out_count);

#ifdef __cplusplus
}
#endif
#endif
