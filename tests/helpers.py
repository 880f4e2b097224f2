"""Shared test helpers: tiny corpora from the reference's own tests, tie-aware comparisons."""
import numpy as np

from oramacore_b200.hostindex import HostStringIndex
from oramacore_b200.types import FieldPostings, StringIndexData


def build_index(docs, fields=("text",)):
    """docs: list of (doc_id, {field: text})."""
    h = HostStringIndex(fields)
    for d, doc in docs:
        h.insert(d, doc)
    h.commit()
    return h


def two_field_golden_index():
    """bm25.rs:912-983: one doc, the term in `title` (tf 2, len 10, avg 8) and `content`
    (tf 1, len 200, avg 150); N=100, df=10 is emulated with 10 docs holding the term and
    document_count=100."""
    n = 10
    rows = np.arange(n, dtype=np.uint32)
    title = FieldPostings(8.0, np.asarray([0, n], np.uint64), rows, np.full(n, 2, np.uint16), np.full(n, 10, np.uint16))
    content = FieldPostings(150.0, np.asarray([0, n], np.uint64), rows, np.full(n, 1, np.uint16), np.full(n, 200, np.uint16))
    return StringIndexData([title, content], n, 100, None)


def assert_topk_equal(got_docs, got_scores, exp_docs, exp_scores, atol=1e-5, tie_eps=1e-6):
    """Scores equal within atol position by position; doc-id sets equal except that
    documents whose score ties (within tie_eps) with the boundary score may swap."""
    got_docs, exp_docs = np.asarray(got_docs), np.asarray(exp_docs)
    got_scores, exp_scores = np.asarray(got_scores, np.float64), np.asarray(exp_scores, np.float64)
    assert got_docs.shape == exp_docs.shape, (got_docs, exp_docs)
    if got_docs.size == 0:
        return
    assert np.allclose(got_scores, exp_scores, rtol=0, atol=atol), (got_scores, exp_scores)
    g, e = set(got_docs.tolist()), set(exp_docs.tolist())
    if g == e:
        return
    boundary = exp_scores[-1]
    for d in g ^ e:
        s = got_scores[got_docs.tolist().index(d)] if d in g else exp_scores[exp_docs.tolist().index(d)]
        assert abs(s - boundary) <= max(tie_eps, atol), f"doc {d} differs and is not a boundary tie ({s} vs {boundary})"
