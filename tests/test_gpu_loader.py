"""The reference's write-operation stream (Index / IndexEmbedding / DeleteDocuments, read/index/mod.rs:1436-1705)
through IndexLoader -> native dictionary -> C ABI stores, then searched end to end with real query strings.
Pins taken from the reference's own integration tests (src/tests/fulltext_search.rs, facets.rs, delete_doc.rs)."""
import numpy as np
import pytest

import oramacore_b200 as ob
from oramacore_b200.hostindex import tokenize
from oramacore_b200.loader import IndexLoader
from oramacore_b200.types import MODE_FULLTEXT, MODE_HYBRID

pytestmark = pytest.mark.gpu


def _index_op(doc_id, text=None, **filters):
    vals = []
    if text is not None:
        toks = tokenize(text)
        terms = {}
        for i, t in enumerate(toks):
            terms.setdefault(t, {"exact_positions": [], "positions": []})["exact_positions"].append(i)
        vals.append({"type": "ScoreString2", "field": "text", "field_length": len(toks), "terms": terms})
    for k, v in filters.items():
        vals.append({"type": {bool: "FilterBool", str: "FilterString"}.get(type(v), "FilterNumber"), "field": k, "value": v})
    return {"type": "Index", "doc_id": doc_id, "indexed_values": vals}


def test_op_stream_fulltext_facets_delete_commit(gpu_ctx):
    ld = IndexLoader(gpu_ctx, ["text"], bool_fields=["bool"], number_fields=["number"], string_filter_fields=["category"])
    # facets.rs:253-342 + a category field
    ld.apply_all([_index_op(1, "text", bool=True, number=1, category="A"),
                  _index_op(2, "text text", bool=False, number=2, category="B"),
                  _index_op(3, "another", bool=True, number=1, category="A")])
    ld.commit()
    tsc = ld.context()
    q = ld.resolve(["text"])
    hits = tsc.execute_batch(ob.TokenScoreParams(mode=MODE_FULLTEXT), q)[0]
    assert sorted(hits.doc_ids.tolist()) == [1, 2] and hits.count == 2
    f = ob.search_facets(tsc, ld.facets, ob.TokenScoreParams(mode=MODE_FULLTEXT),
                         {"bool": {"true": True, "false": True}, "number": {"ranges": [{"from": 0, "to": 10}]}, "category": {}}, texts=q)[0]
    assert f["bool"]["values"] == {"true": 1, "false": 1}
    assert f["number"]["values"] == {"0-10": 2}
    assert f["category"] == {"count": 2, "values": {"A": 1, "B": 1}}
    # prefix expansion (fulltext_search.rs:633-644) and tolerance (:956-1018) through the native dictionary
    ld.apply_all([_index_op(4, "Christopher Nolan"), _index_op(5, "Main Street")])
    assert tsc.execute_batch(ob.TokenScoreParams(mode=MODE_FULLTEXT), ld.resolve(["christoph"]))[0].count == 0   # not committed yet
    ld.commit()
    assert tsc.execute_batch(ob.TokenScoreParams(mode=MODE_FULLTEXT), ld.resolve(["christoph"]))[0].doc_ids.tolist() == [4]
    assert tsc.execute_batch(ob.TokenScoreParams(mode=MODE_FULLTEXT), ld.resolve(["Mxin"], tolerance=1))[0].doc_ids.tolist() == [5]
    assert tsc.execute_batch(ob.TokenScoreParams(mode=MODE_FULLTEXT), ld.resolve(["christoph"], exact=True))[0].count == 0
    # DeleteDocuments: excluded at once, dropped by the next commit; N of the idf follows Index::document_count
    ld.apply({"type": "DeleteDocuments", "doc_ids": [2]})
    assert tsc.execute_batch(ob.TokenScoreParams(mode=MODE_FULLTEXT), q)[0].doc_ids.tolist() == [1]
    v_before = ld.strs.info()["version"]
    ld.commit()
    info = ld.strs.info()
    assert info["version"] == v_before + 1 and info["total_documents"] == 4 and ld.document_count == 4
    assert tsc.execute_batch(ob.TokenScoreParams(mode=MODE_FULLTEXT), q)[0].doc_ids.tolist() == [1]
    ld.close()


def test_shorter_document_ranks_first_and_hybrid_over_the_op_stream(gpu_ctx):
    # fulltext_search.rs:192-251: 100 docs "text " x (i+1): top-5 ids 99..95?  No: BM25 length normalisation favours the
    # SHORT documents for equal tf/len ratio...  the reference pins ids 99..95 with strictly decreasing scores because tf grows
    # with the length; reproduce exactly that ordering.
    dim = 64
    rng = np.random.default_rng(0)
    vecs = rng.standard_normal((100, dim)).astype(np.float32)
    ld = IndexLoader(gpu_ctx, ["text"], embedding_dim=dim)
    ld.apply_all([_index_op(i, "text " * (i + 1)) for i in range(100)])
    ld.apply({"type": "IndexEmbedding", "data": [(i, [vecs[i]]) for i in range(100)]})
    ld.commit()
    tsc = ld.context()
    hits = tsc.execute_batch(ob.TokenScoreParams(mode=MODE_FULLTEXT, limit_hint=5), ld.resolve(["text"]))[0]
    assert hits.doc_ids.tolist() == [99, 98, 97, 96, 95] and np.all(np.diff(hits.scores) < 0) and hits.count == 100
    h = tsc.execute_batch(ob.TokenScoreParams(mode=MODE_HYBRID, similarity=0.0), ld.resolve(["text"]), vecs[7:8])[0]
    assert h.count == 100 and 7 in h.doc_ids.tolist()          # the planted vector hit is fused into the fulltext map
    ld.close()
