"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, the committed golden vectors, the reference's behavioural pins, edge cases, and
size-independent properties at larger sizes.

Bar (BASELINE.json north_star): bit-exact doc-id sets / counts for integer work; cosine and
BM25 scores within 1e-5 fp32 (BM25 is in fact bit-identical by construction: same op order,
no FMA contraction, idf from the same libm)."""
import json
import os

import numpy as np
import pytest

import oramacore_b200 as ob
from helpers import assert_topk_equal, build_index, two_field_golden_index
from oramacore_b200 import synth
from oramacore_b200.types import MODE_FULLTEXT, MODE_HYBRID, MODE_VECTOR, TextQuery

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bm25_known_answers.json")))
ATOL = 1e-5


def _oracle_batch(orc, data, rows, mode, texts=None, qv=None, **kw):
    ix = orc.StrIndex(data) if data is not None else None
    st = rows if (rows is None or isinstance(rows, orc.EmbStore)) else orc.EmbStore(rows)
    sb = orc.SearchBatch(ix, st)
    B = len(texts) if texts is not None else len(qv)
    for i in range(B):
        sb.add(mode, q_vec=None if qv is None else qv[i], text=None if texts is None else texts[i], **kw)
    return sb.run(4)


def _check(hits, ref, atol=ATOL, exact_scores=False):
    od, os_, on, oc = ref
    for i, h in enumerate(hits):
        assert h.count == int(oc[i]), (i, h.count, int(oc[i]))
        n = int(on[i])
        assert len(h.doc_ids) == n, (i, len(h.doc_ids), n)
        if exact_scores:
            assert np.array_equal(h.scores, os_[i, :n]), (i, h.scores, os_[i, :n])
        assert_topk_equal(h.doc_ids, h.scores, od[i, :n], os_[i, :n], atol=atol)


# ------------------------------------------------------------------ vectors
@pytest.mark.parametrize("n,dim,model", [(5000, 768, "BGEBase"), (3001, 384, "BGESmall"), (2000, 1024, "BGELarge"),
                                         (777, 768, "MultilingualE5Base")])
def test_vector_search_parity(gpu_ctx, orc, n, dim, model):
    rows = synth.make_vectors(n, dim, seed=n)
    qv, _ = synth.make_vector_queries(rows, 7, seed=n + 1)
    emb = ob.EmbeddingFieldStorage(gpu_ctx, model)
    emb.insert_batch(np.arange(n, dtype=np.uint64), rows)
    e5 = model.startswith("MultilingualE5")
    st = orc.EmbStore(rows, is_e5=e5)
    for sim in (0.0, 0.05):
        docs, scores, counts = emb.search_batch(qv, 10, sim)
        for i in range(qv.shape[0]):
            ed, es = orc.vector(st, qv[i], 10, sim)
            order = np.argsort(-es, kind="stable")
            assert counts[i] == len(ed)
            assert_topk_equal(docs[i, :counts[i]], scores[i, :counts[i]], ed[order], es[order], atol=ATOL)
    emb.close()


def test_vector_recall_vs_fp64(gpu_ctx, orc):
    n, dim = 20000, 768
    rows = synth.make_vectors(n, dim, seed=5)
    qv, j = synth.make_vector_queries(rows, 16, seed=6)
    emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGEBase")
    emb.insert_batch(np.arange(n, dtype=np.uint64), rows)
    docs, scores, counts = emb.search_batch(qv, 10, -1.0)
    st = orc.EmbStore(rows)
    hit = tot = 0
    for i in range(16):
        ed, ec = orc.vector_f64(st, qv[i], 10)
        assert docs[i, 0] == j[i]
        got = set(docs[i, :counts[i]].tolist())
        for d, c in zip(ed, ec):
            tot += 1
            hit += (int(d) in got) or abs(c - ec[-1]) <= 1e-6
        assert np.allclose(np.sort(scores[i])[::-1], ec, atol=ATOL)
    assert hit / tot >= 0.99
    emb.close()


def test_vector_edge_cases(gpu_ctx, orc):
    emb = ob.EmbeddingFieldStorage(gpu_ctx, dim=64)
    q = np.ones((1, 64), np.float32)
    docs, scores, counts = emb.search_batch(q, 5, 0.0)          # empty store
    assert counts[0] == 0
    rows = synth.make_vectors(3, 64, seed=1)
    emb.insert_batch(np.asarray([7, 8, 9], np.uint64), rows)     # limit > n
    docs, scores, counts = emb.search_batch(rows[1:2], 5, -1.0)
    assert counts[0] == 3 and docs[0, 0] == 8 and abs(scores[0, 0] - 1.0) < 1e-5
    emb.delete(8)                                                # delete (embedding_field.rs:240-242)
    docs, scores, counts = emb.search_batch(rows[1:2], 5, -1.0)
    assert counts[0] == 2 and 8 not in docs[0, :2].tolist()
    assert emb.info()["num_embeddings"] == 2
    emb.insert(5, [rows[0], rows[0]])                            # two chunks of one doc (:232-237)
    out = {}
    emb.search(ob.VectorSearchParams(rows[0], 0.5, 10), out)
    assert abs(out[5] - 2.0) < 1e-5 and abs(out[7] - 1.0) < 1e-5
    z = np.zeros((1, 64), np.float32)                            # zero query -> cos 0 everywhere
    docs, scores, counts = emb.search_batch(z, 2, -1.0)
    assert counts[0] == 2 and np.all(scores[0, :2] == 0.0)
    with pytest.raises(ob.OcError):
        emb.search_batch(q, 5000, 0.0)                           # limit > OC_MAX_TOPK
    emb.close()


def test_vector_filter_and_growth(gpu_ctx, orc):
    n, dim = 4000, 384
    rows = synth.make_vectors(n, dim, seed=9)
    emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGESmall")
    ids = np.arange(n, dtype=np.uint64) * 3 + 1                  # sparse doc ids
    for i in range(0, n, 1000):                                  # incremental inserts force regrowth
        emb.insert_batch(ids[i:i + 1000], rows[i:i + 1000])
    allowed = ids[::7]
    nbits = int(ids.max()) + 1
    fb = orc.make_filter_bits(allowed.tolist(), nbits)
    qv, _ = synth.make_vector_queries(rows, 3, seed=10)
    docs, scores, counts = emb.search_batch(qv, 10, -1.0, fb, nbits)
    st = orc.EmbStore(rows, row_doc_ids=ids)
    for i in range(3):
        ed, es = orc.vector(st, qv[i], 10, -1.0, fb, nbits)
        order = np.argsort(-es, kind="stable")
        assert set(docs[i, :counts[i]].tolist()) <= set(allowed.tolist())
        assert_topk_equal(docs[i, :counts[i]], scores[i, :counts[i]], ed[order], es[order])
    emb.close()


def test_large_limit(gpu_ctx, orc):
    n, dim = 6000, 384
    rows = synth.make_vectors(n, dim, seed=21)
    emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGESmall")
    emb.insert_batch(np.arange(n, dtype=np.uint64), rows)
    qv, _ = synth.make_vector_queries(rows, 2, seed=22)
    st = orc.EmbStore(rows)
    for limit in (100, 1000):
        docs, scores, counts = emb.search_batch(qv, limit, -1.0)
        for i in range(2):
            ed, es = orc.vector(st, qv[i], limit, -1.0)
            order = np.argsort(-es, kind="stable")
            assert_topk_equal(docs[i, :counts[i]], scores[i, :counts[i]], ed[order], es[order])
    emb.close()


# ------------------------------------------------------------------ full text
def test_golden_known_answers_on_gpu(gpu_ctx):
    # bm25.rs:912-983 through the C ABI: multi-field token, weights 2 / 1
    c = G["canonical_two_fields"]
    strs = ob.StringFieldStorage(gpu_ctx, two_field_golden_index())
    q = TextQuery.from_tokens([[(0, 0, 2.0), (1, 0, 1.0)]])
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=[q], limit=10)[0]
    assert hits.count == 10 and abs(float(hits.scores[0]) - c["expected"]) <= c["tol"]
    # bm25.rs:534-563: tf 5, len = avg = 100, N = 100, df = 10
    from oramacore_b200.types import FieldPostings, StringIndexData
    f = FieldPostings(100.0, np.asarray([0, 10], np.uint64), np.arange(10, dtype=np.uint32),
                      np.full(10, 5, np.uint16), np.full(10, 100, np.uint16))
    s2 = ob.StringFieldStorage(gpu_ctx, StringIndexData([f], 10, 100, None))
    h2 = ob.search(gpu_ctx, None, s2, "fulltext", texts=[TextQuery.single_terms([0])], limit=3)[0]
    assert abs(float(h2.scores[0]) - G["scorer_basic"]["expected"]) <= G["scorer_basic"]["tol"]
    strs.close(); s2.close()


def _gpu_ft(ctx, h, term, limit=10, threshold=None, **kw):
    strs = ob.StringFieldStorage(ctx, h.data)
    hits = ob.search(ctx, None, strs, "fulltext", texts=[h.resolve(term, **kw)], limit=limit, threshold=threshold)[0]
    strs.close()
    return hits


def test_reference_behaviour_pins_on_gpu(gpu_ctx):
    h = build_index([(1, {"text": "This is a long text with a lot of words"}), (2, {"text": "This is a smaller text"})])
    r = _gpu_ft(gpu_ctx, h, "text")                       # fulltext_search.rs:146-189
    assert r.count == 2 and r.doc_ids.tolist() == [2, 1] and r.scores[0] > r.scores[1]
    h = build_index([(i, {"text": "text " * (i + 1)}) for i in range(100)])
    r = _gpu_ft(gpu_ctx, h, "text", limit=10)             # :192-251
    assert r.count == 100 and r.doc_ids[:5].tolist() == [99, 98, 97, 96, 95]
    assert all(a > b for a, b in zip(r.scores, r.scores[1:5]))
    h = build_index([(1, {"text": "The pen is on the table"}), (2, {"text": "the pen", "text2": "is on the table"}),
                     (3, {"text": "the pen"})], fields=("text", "text2"))
    assert len(_gpu_ft(gpu_ctx, h, "the pen is on the table", threshold=0.7).doc_ids) == 2   # :478-600
    assert len(_gpu_ft(gpu_ctx, h, "the pen is on the table", threshold=1.0).doc_ids) == 2
    assert len(_gpu_ft(gpu_ctx, h, "pen", threshold=0.0).doc_ids) == 3
    assert len(_gpu_ft(gpu_ctx, h, "pen", threshold=1.0).doc_ids) == 3
    h = build_index([(i, {"text": f"word{i} common"}) for i in range(7)])
    r = _gpu_ft(gpu_ctx, h, "")                           # :890-953 empty term => all docs
    assert r.count == 7 and len(r.doc_ids) == 7
    h = build_index([(1, {"text": "serve the dish"}), (2, {"text": "server the dish"})])
    r = _gpu_ft(gpu_ctx, h, "serve")                      # boost_integration.rs:449-490
    assert r.count == 2 and r.doc_ids[0] == 1
    assert _gpu_ft(gpu_ctx, h, "serve", exact=True).doc_ids.tolist() == [1]


@pytest.mark.parametrize("n_docs,vocab,threshold", [(50000, 3000, None), (50000, 3000, 1.0), (40000, 500, 0.5),
                                                     (16384, 2000, None), (16385, 2000, None)])
def test_fulltext_parity(gpu_ctx, orc, n_docs, vocab, threshold):
    data = synth.make_text_corpus(n_docs, vocab, seed=n_docs + vocab)
    texts = synth.make_text_queries(vocab, 24, seed=7)
    strs = ob.StringFieldStorage(gpu_ctx, data)
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=texts, limit=10, threshold=threshold)
    _check(hits, _oracle_batch(orc, data, None, 0, texts=texts, limit=10, threshold=threshold), exact_scores=True)
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=texts, limit=7, offset=5, threshold=threshold)
    _check(hits, _oracle_batch(orc, data, None, 0, texts=texts, limit=7, offset=5, threshold=threshold),
           exact_scores=True)
    strs.close()


def test_fulltext_multiterm_filter_delete_omc(gpu_ctx, orc):
    n_docs, vocab = 40000, 1500
    data = synth.make_text_corpus(n_docs, vocab, seed=3)
    rng = np.random.default_rng(0)
    # tokens expanding to several index terms with different weights (prefix / fuzzy expansion shape)
    texts = []
    for _ in range(12):
        toks = []
        for _t in range(int(rng.integers(1, 4))):
            k = int(rng.integers(1, 5))
            ids = rng.choice(vocab // 4, size=k, replace=False)
            toks.append([(0, int(t), float(w)) for t, w in zip(ids, rng.choice([1.0, 2.0, 0.5], size=k))])
        texts.append(TextQuery.from_tokens(toks))
    strs = ob.StringFieldStorage(gpu_ctx, data)
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=texts, limit=10)
    _check(hits, _oracle_batch(orc, data, None, 0, texts=texts, limit=10), exact_scores=True)
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=texts, limit=10, threshold=0.6)
    _check(hits, _oracle_batch(orc, data, None, 0, texts=texts, limit=10, threshold=0.6), exact_scores=True)
    # filter (df is counted over filtered docs: collect_contributions_with_filter)
    allowed = np.flatnonzero(rng.random(n_docs) < 0.3)
    fb = orc.make_filter_bits(allowed.tolist(), n_docs)
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=texts, limit=10, filtered_doc_ids=fb, filter_nbits=n_docs)
    _check(hits, _oracle_batch(orc, data, None, 0, texts=texts, limit=10, filter_bits=fb, filter_nbits=n_docs),
           exact_scores=True)
    for h in hits:
        assert set(h.doc_ids.tolist()) <= set(allowed.tolist())
    # OMC multipliers (omc_test.rs: x2 / x3 / x0.5)
    omc_doc = np.sort(rng.choice(n_docs, size=500, replace=False)).astype(np.uint64)
    omc_mult = rng.choice([2.0, 3.0, 0.5], size=500).astype(np.float32)
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=texts, limit=10, omc_doc_ids=omc_doc, omc_mult=omc_mult)
    _check(hits, _oracle_batch(orc, data, None, 0, texts=texts, limit=10, omc_doc=omc_doc, omc_mult=omc_mult),
           exact_scores=True)
    # delete == filter out the doc (string_field.rs:180-182)
    gone = [int(hits[0].doc_ids[0]), int(hits[1].doc_ids[0])]
    for d in gone:
        strs.delete(d)
    keep = orc.make_filter_bits([d for d in range(n_docs) if d not in gone], n_docs)
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=texts, limit=10)
    _check(hits, _oracle_batch(orc, data, None, 0, texts=texts, limit=10, filter_bits=keep, filter_nbits=n_docs),
           exact_scores=True)
    strs.close()


def test_fulltext_sparse_doc_ids_and_unknown_terms(gpu_ctx, orc):
    data = synth.make_text_corpus(20000, 800, seed=8)
    data.row_doc_ids = (np.arange(20000, dtype=np.uint64) * 5 + 3)
    strs = ob.StringFieldStorage(gpu_ctx, data)
    texts = synth.make_text_queries(800, 6, seed=9)
    texts.append(TextQuery.single_terms([799, 5000]))      # unknown term id -> no postings
    texts.append(TextQuery.from_tokens([[]]))              # a token that expands to nothing
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=texts, limit=10)
    _check(hits, _oracle_batch(orc, data, None, 0, texts=texts, limit=10), exact_scores=True)
    assert hits[-1].count == 0 and len(hits[-1].doc_ids) == 0
    strs.close()


# ------------------------------------------------------------------ hybrid + modes
def test_hybrid_parity(gpu_ctx, orc):
    n, dim, vocab, B = 30000, 768, 4000, 16
    rows = synth.make_vectors(n, dim, seed=31)
    qv, _ = synth.make_vector_queries(rows, B, seed=32)
    data = synth.make_text_corpus(n, vocab, seed=33)
    texts = synth.make_text_queries(vocab, B, seed=34)
    emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGEBase")
    emb.insert_batch(np.arange(n, dtype=np.uint64), rows)
    strs = ob.StringFieldStorage(gpu_ctx, data)
    st = orc.EmbStore(rows)
    for kw in (dict(limit=10, similarity=0.0), dict(limit=10, similarity=0.7), dict(limit=5, offset=3, similarity=0.0),
               dict(limit=10, similarity=0.0, threshold=1.0)):
        hits = ob.search(gpu_ctx, emb, strs, "hybrid", texts=texts, q_vecs=qv, **kw)
        _check(hits, _oracle_batch(orc, data, st, 2, texts=texts, qv=qv, **kw))
    # OMC + filter in hybrid
    rng = np.random.default_rng(1)
    omc_doc = np.sort(rng.choice(n, size=2000, replace=False)).astype(np.uint64)
    omc_mult = rng.choice([2.0, 3.0, 0.5], size=2000).astype(np.float32)
    allowed = np.flatnonzero(rng.random(n) < 0.5)
    fb = orc.make_filter_bits(allowed.tolist(), n)
    hits = ob.search(gpu_ctx, emb, strs, "hybrid", texts=texts, q_vecs=qv, limit=10, similarity=0.0,
                     omc_doc_ids=omc_doc, omc_mult=omc_mult, filtered_doc_ids=fb, filter_nbits=n)
    _check(hits, _oracle_batch(orc, data, st, 2, texts=texts, qv=qv, limit=10, similarity=0.0, omc_doc=omc_doc,
                               omc_mult=omc_mult, filter_bits=fb, filter_nbits=n))
    # vector mode through search()
    hits = ob.search(gpu_ctx, emb, None, "vector", q_vecs=qv, limit=10, similarity=0.0, omc_doc_ids=omc_doc,
                     omc_mult=omc_mult)
    _check(hits, _oracle_batch(orc, None, st, 1, qv=qv, limit=10, similarity=0.0, omc_doc=omc_doc, omc_mult=omc_mult))
    emb.close(); strs.close()


def test_hybrid_degenerate_normalisation(gpu_ctx, orc):
    # no fulltext match and no vector hit above the threshold: max == min == 0 -> NaN -> dropped, count kept
    rows = synth.make_vectors(200, 384, seed=1)
    data = synth.make_text_corpus(200, 50, seed=2)
    emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGESmall")
    emb.insert_batch(np.arange(200, dtype=np.uint64), rows)
    strs = ob.StringFieldStorage(gpu_ctx, data)
    texts = [TextQuery.single_terms([4999])]
    qv = -rows[:1]
    hits = ob.search(gpu_ctx, emb, strs, "hybrid", texts=texts, q_vecs=qv, limit=5, similarity=0.99)
    _check(hits, _oracle_batch(orc, data, orc.EmbStore(rows), 2, texts=texts, qv=qv, limit=5, similarity=0.99))
    assert hits[0].count == 0
    emb.close(); strs.close()


def test_reference_shaped_execute(gpu_ctx, orc):
    # TokenScoreContext::execute(&params, &mut HashMap) shape (token_score.rs:460-464)
    h = build_index([(i, {"text": "text " * (i + 1)}) for i in range(20)])
    strs = ob.StringFieldStorage(gpu_ctx, h.data)
    tsc = ob.TokenScoreContext(gpu_ctx, None, strs)
    res = {}
    count = tsc.execute(ob.TokenScoreParams(mode=MODE_FULLTEXT, limit_hint=20), res, text=h.resolve("text"))
    assert count == 20 and len(res) == 20 and max(res, key=res.get) == 19
    strs.close()


# ------------------------------------------------------------------ size-independent properties at scale
def test_scale_properties(gpu_ctx, orc):
    n, dim = 300000, 768
    rows = synth.make_vectors(n, dim, seed=77)
    qv, j = synth.make_vector_queries(rows, 9, seed=78)
    emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGEBase")
    emb.insert_batch(np.arange(n, dtype=np.uint64), rows)
    docs, scores, counts = emb.search_batch(qv, 10, -1.0)
    assert np.all(counts == 10) and np.all(docs[:, 0] == j)                   # planted neighbour is rank 1
    assert np.all(np.diff(scores, axis=1) <= 0)                               # sortedness
    d2, s2, _ = emb.search_batch(qv[::-1].copy(), 10, -1.0)                   # batch-order independence
    assert np.array_equal(d2[::-1], docs) and np.array_equal(s2[::-1], scores)
    d1, s1, _ = emb.search_batch(qv[:1], 10, -1.0)                            # QB=1 path == QB=4 path
    assert np.array_equal(d1[0], docs[0]) and np.array_equal(s1[0], scores[0])
    # each returned score equals an independent fp64 cosine of that row
    for i in range(3):
        x = rows[docs[i].astype(np.int64)].astype(np.float64)
        q = qv[i].astype(np.float64)
        c = x @ q / (np.linalg.norm(x, axis=1) * np.linalg.norm(q))
        assert np.allclose(c, scores[i], atol=ATOL)
    # idempotence
    d3, s3, _ = emb.search_batch(qv, 10, -1.0)
    assert np.array_equal(d3, docs) and np.array_equal(s3, scores)
    emb.close()
    # fulltext at 1M docs: counts equal the oracle's, top-10 bit-identical
    data = synth.make_text_corpus(1000000, 50000, seed=79)
    texts = synth.make_text_queries(50000, 8, seed=80)
    strs = ob.StringFieldStorage(gpu_ctx, data)
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=texts, limit=10)
    _check(hits, _oracle_batch(orc, data, None, 0, texts=texts, limit=10), exact_scores=True)
    strs.close()


def test_incremental_insert_commit_delete(gpu_ctx, orc):
    # StringFieldStorage::insert / delete / compact (string_field.rs:155-191) through oc_str_insert/_commit
    rng = np.random.default_rng(3)
    vocab = [f"w{i}" for i in range(300)]
    docs = {}
    for d in range(0, 4000, 2):                      # sparse, even doc ids
        n = int(rng.integers(3, 40))
        docs[d] = " ".join(rng.choice(vocab, size=n, p=None))
    from oramacore_b200.hostindex import HostStringIndex, tokenize
    tid = {t: i for i, t in enumerate(sorted(vocab))}

    def feed(strs, items):
        for d, text in items:
            toks = tokenize(text)
            counts = {}
            for t in toks:
                counts[tid[t]] = counts.get(tid[t], 0) + 1
            strs.insert(d, 0, len(toks), counts)

    strs = ob.StringFieldStorage.empty(gpu_ctx, 1)
    items = sorted(docs.items())
    feed(strs, items[:1500])
    strs.commit()
    feed(strs, items[1500:])                         # second batch + a replacement + deletes
    docs[10] = "w1 w1 w1 w2"
    feed(strs, [(10, docs[10])])
    strs.commit()
    for gone in (20, 30):
        strs.delete(gone)
        docs.pop(gone)
    strs.commit()

    h = HostStringIndex(["text"])
    for d, text in sorted(docs.items()):
        h.insert(d, {"text": text})
    h.commit()
    # the incremental store numbers terms by the fixed vocabulary; rebuild the oracle's view with the same ids
    assert h.terms[0] == sorted(set(t for text in docs.values() for t in tokenize(text)))
    remap = {i: tid[t] for i, t in enumerate(h.terms[0])}
    texts = []
    for _ in range(10):
        ids = rng.choice(len(h.terms[0]), size=3, replace=False)
        texts.append(([int(i) for i in ids], [remap[int(i)] for i in ids]))
    ix = orc.StrIndex(h.data)
    sb = orc.SearchBatch(ix, None)
    for o_ids, _ in texts:
        sb.add(0, limit=10, text=TextQuery.single_terms(o_ids))
    od, os_, on, oc = sb.run(2)
    hits = ob.search(gpu_ctx, None, strs, "fulltext", texts=[TextQuery.single_terms(g) for _, g in texts], limit=10)
    for i, hh in enumerate(hits):
        assert hh.count == int(oc[i])
        assert np.array_equal(hh.scores, os_[i, :on[i]]), (hh.scores, os_[i, :on[i]])
        assert_topk_equal(hh.doc_ids, hh.scores, od[i, :on[i]], os_[i, :on[i]])
    assert strs.info()["total_documents"] == len(docs)
    strs.close()


def test_bm25_shared_term_precompute_is_bit_identical(gpu_ctx, orc):
    # the batch-level sharing of per-posting contributions must not change a single bit
    data = synth.make_text_corpus(60000, 2000, seed=17)
    texts = synth.make_text_queries(2000, 64, seed=18)
    strs = ob.StringFieldStorage(gpu_ctx, data)
    out = {}
    for mode in ("off", "force"):
        os.environ["OC_BM25_SHARE"] = mode
        try:
            out[mode] = ob.search(gpu_ctx, None, strs, "fulltext", texts=texts, limit=10, threshold=0.5)
        finally:
            os.environ.pop("OC_BM25_SHARE", None)
    for a, b in zip(out["off"], out["force"]):
        assert a.count == b.count and np.array_equal(a.doc_ids, b.doc_ids) and np.array_equal(a.scores, b.scores)
    _check(out["force"], _oracle_batch(orc, data, None, 0, texts=texts, limit=10, threshold=0.5), exact_scores=True)
    strs.close()


def test_concurrent_callers_share_one_ctx(gpu_ctx, orc):
    # many searches run concurrently in the reference (tokio workers, read/mod.rs:621); handles are
    # Send+Sync here: calls on one ctx serialise internally and must not corrupt each other
    import threading
    n, dim = 20000, 384
    rows = synth.make_vectors(n, dim, seed=51)
    data = synth.make_text_corpus(n, 1500, seed=52)
    emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGESmall")
    emb.insert_batch(np.arange(n, dtype=np.uint64), rows)
    strs = ob.StringFieldStorage(gpu_ctx, data)
    qv, _ = synth.make_vector_queries(rows, 12, seed=53)
    qv_p = ob.pinned_empty(qv.shape)                       # pinned inputs take the direct-DMA path
    qv_p[...] = qv
    texts = synth.make_text_queries(1500, 12, seed=54)
    ref = _oracle_batch(orc, data, orc.EmbStore(rows), 2, texts=texts, qv=qv, limit=10, similarity=0.0)
    errs = []

    def worker(k):
        try:
            for _ in range(5):
                hits = ob.search(gpu_ctx, emb, strs, "hybrid", texts=texts, q_vecs=qv_p if k % 2 else qv, limit=10, similarity=0.0)
                _check(hits, ref)
        except Exception as e:  # noqa
            errs.append(e)

    th = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    emb.close(); strs.close()


def test_batcher_coalesces_single_query_callers(gpu_ctx, orc):
    """oc_batcher_*: 8 threads submit one hybrid query each (the reference's one-search-per-task
    shape); results must equal the direct batched oc_search and calls must be coalesced."""
    import threading
    n, dim, vocab, B = 30000, 384, 2000, 48
    rows = synth.make_vectors(n, dim, seed=71)
    qv, _ = synth.make_vector_queries(rows, B, seed=72)
    data = synth.make_text_corpus(n, vocab, seed=73)
    texts = synth.make_text_queries(vocab, B, seed=74)
    emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGESmall")
    emb.insert_batch(np.arange(n, dtype=np.uint64), rows)
    strs = ob.StringFieldStorage(gpu_ctx, data)
    tsc = ob.TokenScoreContext(gpu_ctx, emb, strs)
    params = ob.TokenScoreParams(mode=ob.MODE_HYBRID, limit_hint=10, similarity=0.0)
    direct = tsc.execute_batch(params, texts, qv)
    bat = ob.SearchBatcher(tsc, max_batch=16, max_wait_us=20000)
    got, errs = [None] * B, []

    def worker(t):
        try:
            for i in range(t, B, 8):
                got[i] = bat.search(params, texts[i], qv[i])
        except Exception as e:   # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for i in range(B):
        assert got[i].count == direct[i].count
        # a query's result does not depend on which other queries share its batch (bit-identical BM25,
        # exact fp32 re-score of the vector hits)
        assert np.array_equal(got[i].doc_ids, direct[i].doc_ids) and np.array_equal(got[i].scores, direct[i].scores)
    st = bat.stats()
    assert st["queries"] == B and st["batches"] < B, st
    # a filtered query is not coalesced: it goes straight through
    fb = orc.make_filter_bits(list(range(0, n, 2)), n)
    pf = ob.TokenScoreParams(mode=ob.MODE_FULLTEXT, limit_hint=10, filtered_doc_ids=fb, filter_nbits=n)
    h = bat.search(pf, texts[0], None)
    ref = tsc.execute_batch(pf, [texts[0]], None)[0]
    assert h.count == ref.count and np.array_equal(h.doc_ids, ref.doc_ids)
    assert bat.stats()["direct"] == 1
    bat.close(); strs.close(); emb.close()
