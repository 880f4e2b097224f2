"""The oracle (CPU restatement) against the reference's own pins for this path:
closed-form known answers of bm25.rs and the ordering/count pins of src/tests/*.rs."""
import json
import os

import numpy as np
import pytest

from helpers import build_index, two_field_golden_index
from oramacore_b200.types import TextQuery

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bm25_known_answers.json")))
f32 = np.float32


def test_scorer_basic_known_answer(orc):
    c = G["scorer_basic"]  # bm25.rs:534-563
    got = orc.bm25_legacy_add(c["tf"], c["field_len"], c["avg_len"], c["total_docs"], c["df"], c["k"], c["weight"],
                              c["b"], c["boost"])
    assert abs(got - c["expected"]) <= c["tol"]
    # pieces
    assert abs(orc.idf(100.0, 10) - float(np.log1p(f32(90.5) / f32(10.5)))) < 1e-7
    assert orc.normalized_tf(5, 100, 100.0, 0.75) == 5.0


def test_canonical_two_fields_known_answer(orc):
    c = G["canonical_two_fields"]  # bm25.rs:912-983, through the full search_full_text restatement
    data = two_field_golden_index()
    ix = orc.StrIndex(data)
    q = TextQuery.from_tokens([[(0, 0, 2.0), (1, 0, 1.0)]])  # one token, two fields, weights 2 / 1
    docs, scores = orc.fulltext(ix, q)
    assert len(docs) == 10
    assert abs(float(scores[0]) - c["expected"]) <= c["tol"]
    assert abs(orc.normalized_tf(2, 10, 8.0) - c["title_ntf"]) < 1e-6
    assert abs(orc.normalized_tf(1, 200, 150.0) - c["content_ntf"]) < 1e-6


def test_relations(orc):
    add = lambda **k: orc.bm25_legacy_add(k.get("tf", 5), k.get("len", 100), k.get("avg", 100.0), 100.0, 10, 1.2,
                                          k.get("w", 1.0), k.get("b", 0.75), k.get("boost", 1.0))
    assert add(boost=2.0) > add(boost=1.0) > add(boost=0.5)                 # bm25.rs:566-618
    assert add(len=200, b=0.2) > add(len=200, b=0.9)                        # :673-716
    s = [add(w=w) for w in (0.5, 1.0, 1.5, 2.0, 3.0)]                       # :867-909
    assert all(b > a for a, b in zip(s, s[1:]))
    assert 1.0 < s[3] / s[1] < 1.5
    # two fields of one doc (weights 2 and 1) beat the single-field closed form (:619-668)
    single = orc.idf(100.0, 10) * f32(2.2) * f32(5.0) / f32(6.2)
    assert add(w=2.0) + add(w=1.0) > single
    # 2x field weight is a meaningful gain (:715-779: ratio > 1.05)
    assert add(w=2.0) / add(w=1.0) > 1.05
    # title (weight 3, tf 3, len = avg = 50) contributes more than content (weight 1, tf 3, len = avg = 200) (:781-866)
    assert add(tf=3, len=50, avg=50.0, w=3.0) > add(tf=3, len=200, avg=200.0, w=1.0) > 0
    # canonical <= sum of per-field (:986-1043)
    idf = orc.idf(100.0, 10)
    n1, n2 = orc.normalized_tf(3, 50, 40.0), orc.normalized_tf(2, 100, 80.0)
    canonical = orc.bm25f_score(f32(2.0) * f32(n1) + f32(1.0) * f32(n2), 1.2, idf)
    indiv = 2.0 * orc.bm25f_score(n1, 1.2, idf) + 1.0 * orc.bm25f_score(n2, 1.2, idf)
    assert 0 < canonical <= indiv + 1e-6


def test_e5_rescale(orc):
    # python/embeddings.rs:71-92
    assert orc.rescale_score(0.5, False) == 0.5
    assert orc.rescale_score(0.5, True) == 0.0
    assert orc.rescale_score(1.2, True) == 1.0
    assert abs(orc.rescale_score(0.85, True) - (f32(0.85) - f32(0.7)) / (f32(1.0) - f32(0.7))) < 1e-7


def _search_ft(orc, h, term, limit=10, threshold=None, **kw):
    q = h.resolve(term, **kw)
    sb = orc.SearchBatch(orc.StrIndex(h.data), None)
    sb.add(0, limit=limit, text=q, threshold=threshold)
    od, os_, on, oc = sb.run()
    return od[0, :on[0]].tolist(), os_[0, :on[0]].tolist(), int(oc[0])


def test_documents_order(orc):
    # src/tests/fulltext_search.rs:146-189 — the shorter document ranks first
    h = build_index([(1, {"text": "This is a long text with a lot of words"}), (2, {"text": "This is a smaller text"})])
    docs, scores, count = _search_ft(orc, h, "text")
    assert count == 2 and docs == [2, 1] and scores[0] > scores[1]


def test_documents_limit(orc):
    # fulltext_search.rs:192-251 — "text " x (i+1): top ids 99..95, strictly decreasing
    h = build_index([(i, {"text": "text " * (i + 1)}) for i in range(100)])
    docs, scores, count = _search_ft(orc, h, "text", limit=10)
    assert count == 100 and len(docs) == 10
    assert docs[:5] == [99, 98, 97, 96, 95]
    assert all(a > b for a, b in zip(scores, scores[1:5]))


def test_threshold_counts(orc):
    # fulltext_search.rs:478-600
    h = build_index([(1, {"text": "The pen is on the table"}), (2, {"text": "the pen", "text2": "is on the table"}),
                     (3, {"text": "the pen"})], fields=("text", "text2"))
    assert len(_search_ft(orc, h, "the pen is on the table", threshold=0.7)[0]) == 2
    assert len(_search_ft(orc, h, "the pen is on the table", threshold=1.0)[0]) == 2
    assert len(_search_ft(orc, h, "pen", threshold=0.0)[0]) == 3
    assert len(_search_ft(orc, h, "pen", threshold=1.0)[0]) == 3


def test_empty_term_matches_all(orc):
    # fulltext_search.rs:890-953
    h = build_index([(i, {"text": f"word{i} common"}) for i in range(7)])
    docs, _, count = _search_ft(orc, h, "", limit=10)
    assert count == 7 and len(docs) == 7


def test_exact_vs_prefix(orc):
    # fulltext_search.rs:633-644 prefix; boost_integration.rs:449-490 exact outranks prefix
    h = build_index([(1, {"text": "serve the dish"}), (2, {"text": "server the dish"})])
    docs, scores, count = _search_ft(orc, h, "serve")
    assert count == 2 and docs[0] == 1
    docs, _, count = _search_ft(orc, h, "serve", exact=True)
    assert count == 1 and docs == [1]
    docs, _, count = _search_ft(orc, h, "servr", tolerance=1)   # fulltext_search.rs:956-1018
    assert count >= 1


def test_exact_prefix_counts(orc):
    # fulltext_search.rs:603-757 (test_fulltext_exact): originals only when exact, prefix hits otherwise
    h = build_index([(1, {"text": "Christopher Nolan"}), (2, {"text": "Foxes"}), (3, {"text": "Fox"})])
    assert _search_ft(orc, h, "christoph", exact=True)[2] == 0
    assert _search_ft(orc, h, "christoph", exact=False)[2] == 1
    assert _search_ft(orc, h, "Fox", exact=True)[2] == 1
    assert _search_ft(orc, h, "Foxes", exact=True)[2] == 1
    assert _search_ft(orc, h, "Fox", exact=False)[2] == 2


def test_exact_with_threshold_picks_the_right_doc(orc):
    # fulltext_search.rs:758-887: exact "Fox table" matches both docs (shared "table"); threshold 1.0 keeps
    # only the document holding BOTH exact terms
    h = build_index([(1, {"text": "Foxes table"}), (2, {"text": "Fox table"})])
    assert _search_ft(orc, h, "Fox table", exact=True)[2] == 2
    assert _search_ft(orc, h, "Foxes table", exact=True)[2] == 2
    docs, _, count = _search_ft(orc, h, "Fox table", exact=True, threshold=1.0)
    assert count == 1 and docs == [2]
    docs, _, count = _search_ft(orc, h, "Foxes table", exact=True, threshold=1.0)
    assert count == 1 and docs == [1]


def test_field_boost_raises_the_score(orc):
    # src/tests/boost_integration.rs:12-135: "machine learning" over title + content; boosting the field that holds
    # the phrase densely (doc1's title) raises doc1's score by more than 8 %, and more than boosting content does
    docs = [(1, {"title": "machine learning",
                 "content": "This comprehensive document provides detailed information about various algorithms and techniques "
                            "used in modern data processing applications. The field includes machine learning which encompasses "
                            "many different approaches and methodologies for analysis and prediction."}),
            (2, {"title": "data analysis techniques",
                 "content": "Advanced machine learning models and frameworks for comprehensive data analysis, statistical "
                            "processing, and predictive modeling in various business applications and research contexts."}),
            (3, {"title": "statistical methods",
                 "content": "Comprehensive overview of machine learning methodologies combined with traditional statistical "
                            "approaches for effective data analysis, pattern recognition, and business intelligence applications."})]
    h = build_index(docs, fields=("title", "content"))

    def score_of_doc1(boost):
        d, s, count = _search_ft(orc, h, "machine learning", boost=boost, properties=["title", "content"])
        assert count > 0
        return dict(zip(d, s))[1]

    none, title, content = score_of_doc1({}), score_of_doc1({"title": 3.0, "content": 1.0}), score_of_doc1({"title": 1.0, "content": 3.0})
    assert title > none and title > content
    assert title / none > 1.08


def test_vector_contract(orc):
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((500, 64)).astype(np.float32)
    st = orc.EmbStore(rows)
    q = rows[17] + 0.01 * rng.standard_normal(64).astype(np.float32)
    docs, scores = orc.vector(st, q, 5, 0.0)
    od, oc = orc.vector_f64(st, q, 5)
    assert 17 in docs.tolist()
    ref = {int(d): c for d, c in zip(od, oc)}
    for d, s in zip(docs, scores):
        assert abs(ref[int(d)] - float(s)) < 1e-5
    # similarity threshold drops hits (vector_search.rs:11-121: count shrinks under a higher threshold)
    d2, _ = orc.vector(st, q, 5, 0.9)
    assert len(d2) == 1 and d2[0] == 17
    # chunks of one document accumulate (embedding_field.rs:273-274)
    st2 = orc.EmbStore(np.stack([rows[17], rows[17]]), row_doc_ids=np.asarray([5, 5], np.uint64))
    d3, s3 = orc.vector(st2, rows[17], 2, 0.0)
    assert d3.tolist() == [5] and abs(float(s3[0]) - 2.0) < 1e-5


def test_hybrid_combine_and_topn(orc):
    # token_score.rs:393-422: folds start at 0.0, (v-min)/(max-min), fulltext += vector
    vec = (np.asarray([1, 2], np.uint64), np.asarray([0.8, 0.4], np.float32))
    ft = (np.asarray([2, 3], np.uint64), np.asarray([4.0, 2.0], np.float32))
    d, s = orc.hybrid_combine(vec, ft)
    exp = {1: f32(0.8) / f32(4.0), 2: f32(4.0) / f32(4.0) + f32(0.4) / f32(4.0), 3: f32(2.0) / f32(4.0)}
    assert d.tolist() == [1, 2, 3]
    for di, si in zip(d, s):
        assert abs(float(si) - float(exp[int(di)])) < 1e-7
    # max == min => NaN scores, dropped by top_n but still counted (search.rs:482)
    d0, s0 = orc.hybrid_combine((np.zeros(0, np.uint64), np.zeros(0, np.float32)),
                                (np.asarray([9], np.uint64), np.asarray([0.0], np.float32)))
    assert np.isnan(s0[0])
    td, ts = orc.top_n((d0, s0), 5)
    assert len(td) == 0
    # omc (search.rs:39-48; omc_test.rs ratios x2 / x0.5)
    d1, s1 = orc.apply_omc((d, s.copy()), np.asarray([1, 3], np.uint64), np.asarray([2.0, 0.5], np.float32))
    assert abs(s1[0] / s[0] - 2.0) < 1e-6 and abs(s1[2] / s[2] - 0.5) < 1e-6 and s1[1] == s[1]
    # top_n: descending, ties by ascending doc id
    td, ts = orc.top_n((np.asarray([5, 3, 9], np.uint64), np.asarray([1.0, 1.0, 2.0], np.float32)), 2)
    assert td.tolist() == [9, 3]


def test_token_bit_wraps(orc):
    # token_score.rs:293 `1 << term_index` on u32 (release build wraps the shift amount)
    h = build_index([(0, {"text": " ".join(f"w{i}" for i in range(40))})])
    docs, _, count = _search_ft(orc, h, " ".join(f"w{i}" for i in range(40)), threshold=0.5)
    assert count == 1
