"""Facet counts over the score set on the GPU (oc_search_facets) against (1) the reference's own pinned answers
(src/tests/facets.rs: number ranges :10-98, term-restricted counts :253-342, filters ignored :408-460) and (2) a numpy
restatement of FacetContext::execute (read/index/facet.rs:147-209: count the variant's documents that are keys of the
score map) over the oracle's score maps, in fulltext and hybrid mode, identity and sparse document ids."""
import numpy as np
import pytest

import oramacore_b200 as ob
from helpers import build_index
from oramacore_b200 import filters as F
from oramacore_b200 import synth
from oramacore_b200.types import MODE_FULLTEXT, MODE_HYBRID, MODE_VECTOR

pytestmark = pytest.mark.gpu


def _ft(ctx, h):
    return ob.TokenScoreContext(ctx, None, ob.StringFieldStorage(ctx, h.data))


def test_reference_pin_number_ranges(gpu_ctx):
    # facets.rs:10-98: 100 docs "text " x (i+1), number = i, term "text"
    h = build_index([(i, {"text": "text " * (i + 1)}) for i in range(100)])
    tsc = _ft(gpu_ctx, h)
    st = ob.FacetStore(gpu_ctx, 100)
    st.add_number_field("number", np.arange(100), np.arange(100, dtype=np.float64))
    ranges = [(0, 10), (0.5, 10.5), (-10, 10), (-10, -1), (1, 100), (99, 105), (102, 105)]
    r = ob.search_facets(tsc, st, ob.TokenScoreParams(mode=MODE_FULLTEXT), {"number": {"ranges": [{"from": a, "to": b} for a, b in ranges]}},
                         texts=[h.resolve("text")])[0]
    assert r["number"]["count"] == 7
    assert r["number"]["values"] == {"-10--1": 0, "-10-10": 11, "0-10": 11, "0.5-10.5": 10, "1-100": 99, "102-105": 0, "99-105": 1}
    st.close(); tsc.str.close()


def test_reference_pin_counts_follow_the_term_and_ignore_the_filter(gpu_ctx):
    # facets.rs:253-342: the document that does not match the term is not counted
    h = build_index([(1, {"text": "text"}), (2, {"text": "text text"}), (3, {"text": "another"})])
    tsc = _ft(gpu_ctx, h)
    st = ob.FacetStore(gpu_ctx, 4)
    st.add_bool_field("bool", [1, 3], [2])
    st.add_number_field("number", [1, 2, 3], [1.0, 2.0, 1.0])
    r = ob.search_facets(tsc, st, ob.TokenScoreParams(mode=MODE_FULLTEXT),
                         {"bool": {"true": True, "false": True}, "number": {"ranges": [{"from": 0, "to": 10}]}}, texts=[h.resolve("text")])[0]
    assert r["bool"] == {"count": 2, "values": {"true": 1, "false": 1}}
    assert r["number"] == {"count": 1, "values": {"0-10": 2}}
    st.close(); tsc.str.close()
    # facets.rs:408-460: term "", where category = A -> facets still report A: 5, B: 5
    h = build_index([(i, {"title": f"title {i}"}) for i in range(10)], fields=("title",))
    tsc = _ft(gpu_ctx, h)
    st = ob.FacetStore(gpu_ctx, 10)
    st.add_string_field("category", {"A": range(0, 10, 2), "B": range(1, 10, 2)})
    where = F.to_bitmap(F.Ids(range(0, 10, 2)), 10)
    p = ob.TokenScoreParams(mode=MODE_FULLTEXT, filtered_doc_ids=where, filter_nbits=10)
    hits = tsc.execute_batch(p, [h.resolve("")])[0]
    assert hits.count == 5                                                     # the hits ARE filtered
    r = ob.search_facets(tsc, st, p, {"category": {}}, texts=[h.resolve("")])[0]
    assert r == {"category": {"count": 2, "values": {"A": 5, "B": 5}}}          # the facets are not (search.rs:361-396)
    st.close(); tsc.str.close()


def _oracle_counts(keys, variants):
    ks = set(int(k) for k in keys)
    return {label: sum(1 for d in docs if int(d) in ks) for label, docs in variants.items()}


@pytest.mark.parametrize("mode,sparse_ids", [(MODE_FULLTEXT, False), (MODE_FULLTEXT, True), (MODE_HYBRID, False), (MODE_HYBRID, True),
                                              (MODE_VECTOR, False)])
def test_random_corpus_against_the_oracle_score_maps(gpu_ctx, orc, mode, sparse_ids):
    n, dim, vocab, B = 40000, 384, 3000, 12
    rng = np.random.default_rng(9)
    rows = synth.make_vectors(n, dim, seed=61)
    qv, _ = synth.make_vector_queries(rows, B, seed=62)
    data = synth.make_text_corpus(n, vocab, seed=63)
    texts = synth.make_text_queries(vocab, B, seed=64)
    ids = (np.arange(n, dtype=np.uint64) * 3 + 2) if sparse_ids else np.arange(n, dtype=np.uint64)
    if sparse_ids:
        data.row_doc_ids = ids
    nbits = int(ids.max()) + 1
    emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGESmall")
    emb.insert_batch(ids, rows)
    strs = ob.StringFieldStorage(gpu_ctx, data)
    for gone in ids[[5, 77, 4000]].tolist():       # uncommitted deletes stay excluded from the facet score map
        strs.delete(gone); emb.delete(gone)
    deleted = np.zeros(n, np.uint8); deleted[[5, 77, 4000]] = 1
    tsc = ob.TokenScoreContext(gpu_ctx, emb if mode != MODE_FULLTEXT else None, strs if mode != MODE_VECTOR else None)
    flag = rng.random(n) < 0.3
    cat = rng.integers(0, 6, size=n)
    price = np.round(rng.gamma(2.0, 30.0, size=n), 2)
    st = ob.FacetStore(gpu_ctx, nbits)
    st.add_bool_field("in_stock", ids[flag], ids[~flag])
    cat_docs = {f"c{k}": np.concatenate([ids[cat == k], ids[(cat == (k + 1) % 6) & (rng.random(n) < 0.1)]]) for k in range(6)}   # a doc may hold 2 keys
    st.add_string_field("category", cat_docs)
    st.add_number_field("price", ids, price)
    ranges = [(0, 20), (20, 50.5), (50.5, 1e9), (-5, -1)]
    facets = {"in_stock": {"true": True, "false": True}, "category": {}, "price": {"ranges": [{"from": a, "to": b} for a, b in ranges]}}
    where = F.to_bitmap(F.Ids(ids[::2]), nbits)                               # must not influence the counts
    p = ob.TokenScoreParams(mode=mode, similarity=0.0, filtered_doc_ids=where, filter_nbits=nbits)
    got = ob.search_facets(tsc, st, p, facets, texts=texts if mode != MODE_VECTOR else None, q_vecs=qv if mode != MODE_FULLTEXT else None)
    # oracle: the keys of the UNFILTERED score map (deleted docs removed), then plain set counting
    alive = orc.make_filter_bits(ids[deleted == 0].tolist(), nbits)
    ix = orc.StrIndex(data)
    est = orc.EmbStore(rows, row_doc_ids=ids, deleted=deleted)
    variants = {"in_stock": {"true": ids[flag], "false": ids[~flag]}, "category": cat_docs,
                "price": {f"{ob.engine._number_label(a)}-{ob.engine._number_label(b)}": ids[(price >= a) & (price <= b)] for a, b in ranges}}
    for q in range(B):
        if mode == MODE_VECTOR:
            keys = orc.vector(est, qv[q], 10, 0.0)[0]
        else:
            ft = orc.fulltext(ix, texts[q], filter_bits=alive, filter_nbits=nbits)
            keys = ft[0] if mode == MODE_FULLTEXT else orc.hybrid_combine(orc.vector(est, qv[q], 10, 0.0), ft)[0]
        for name in facets:
            exp = _oracle_counts(keys, variants[name])
            assert got[q][name]["values"] == exp, (q, name, got[q][name]["values"], exp)
            assert got[q][name]["count"] == len(exp)
    st.close(); emb.close(); strs.close()
