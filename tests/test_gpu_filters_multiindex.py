"""Device-resident filters (oc_filter_*: FilterResult And / Or / Not evaluated on the GPU, filter.rs:344-392) and the
multi-index union (search_on_indexes, search.rs:304-338) through the C ABI, against filters.py / the oracle."""
import numpy as np
import pytest

import oramacore_b200 as ob
from helpers import assert_topk_equal
from oramacore_b200 import filters as F
from oramacore_b200 import synth
from oramacore_b200.types import MODE_FULLTEXT, MODE_HYBRID, FieldPostings, StringIndexData

pytestmark = pytest.mark.gpu


def test_device_filter_algebra_matches_the_host_tree(gpu_ctx):
    rng = np.random.default_rng(4)
    n = 100_003
    a, b, c = (rng.choice(n, size=k, replace=False) for k in (30000, 45000, 400))
    expr = F.Or(F.And(F.Ids(a), F.Not(F.Ids(b))), F.Ids(c))
    host = F.to_bitmap(expr, n)
    dev = ob.DeviceFilter.from_expr(gpu_ctx, expr, n)
    assert np.array_equal(dev.read(), host)
    assert dev.count() == int(sum(bin(int(w)).count("1") for w in host))
    full = ob.DeviceFilter.from_expr(gpu_ctx, F.Not(F.Ids([])), 70)
    assert full.count() == 70 and int(full.read()[1]) >> 6 == 0          # padding bits stay clear
    # execute_filter's rule: AND(where, NOT(uncommitted deletes))
    ex = ob.DeviceFilter.from_expr(gpu_ctx, F.And(F.Ids(range(0, 200, 2)), F.Not(F.Ids([4, 5]))), 200)
    assert np.array_equal(ex.read(), F.execute_filter(F.Ids(range(0, 200, 2)), [4, 5], 200))
    for f in (dev, full, ex):
        f.close()


def test_search_with_device_filter_equals_host_bitmap(gpu_ctx, orc):
    n, dim, vocab, B = 30000, 384, 2000, 24
    rows = synth.make_vectors(n, dim, seed=41)
    qv, _ = synth.make_vector_queries(rows, B, seed=42)
    data = synth.make_text_corpus(n, vocab, seed=43)
    texts = synth.make_text_queries(vocab, B, seed=44)
    emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGESmall")
    emb.insert_batch(np.arange(n, dtype=np.uint64), rows)
    strs = ob.StringFieldStorage(gpu_ctx, data)
    expr = F.And(F.Ids(range(0, n, 3)), F.Not(F.Ids(range(0, n, 15))))
    bits = F.to_bitmap(expr, n)
    dev = ob.DeviceFilter.from_expr(gpu_ctx, expr, n)
    tsc = ob.TokenScoreContext(gpu_ctx, emb, strs)
    for mode in (MODE_FULLTEXT, MODE_HYBRID):
        a = tsc.execute_batch(ob.TokenScoreParams(mode=mode, similarity=0.0, filtered_doc_ids=bits, filter_nbits=n), texts, qv)
        for _ in range(2):   # the handle is reused across calls, nothing is re-uploaded
            b = tsc.execute_batch(ob.TokenScoreParams(mode=mode, similarity=0.0, device_filter=dev), texts, qv)
            for x, y in zip(a, b):
                assert x.count == y.count and np.array_equal(x.doc_ids, y.doc_ids) and np.array_equal(x.scores, y.scores)
        ix, st = orc.StrIndex(data), orc.EmbStore(rows)
        sb = orc.SearchBatch(ix, st)
        for i in range(B):
            sb.add(mode, limit=10, similarity=0.0, q_vec=qv[i], text=texts[i], filter_bits=bits, filter_nbits=n)
        od, os_, on, oc = sb.run(2)
        for i, h in enumerate(b):
            assert h.count == int(oc[i])
            assert_topk_equal(h.doc_ids, h.scores, od[i, :on[i]], os_[i, :on[i]])
    dev.close(); emb.close(); strs.close()


def _split_text(data, parts):
    """documents d with d % parts == i go to index i (own row space, own avg length / N like separate indexes)."""
    f = data.fields[0]
    df = np.diff(f.term_offsets.astype(np.int64))
    term_of = np.repeat(np.arange(f.n_terms, dtype=np.int64), df)
    out = []
    for i in range(parts):
        docs = np.arange(i, data.n_rows, parts, dtype=np.uint64)
        sel = (f.post_row % parts) == i
        offs = np.zeros(f.n_terms + 1, np.uint64)
        offs[1:] = np.cumsum(np.bincount(term_of[sel], minlength=f.n_terms)).astype(np.uint64)
        lens = np.zeros(data.n_rows, np.int64)
        lens[f.post_row[sel]] = f.post_len[sel]
        avg = float(lens[docs.astype(np.int64)].mean())
        fp = FieldPostings(avg, offs, (f.post_row[sel] // parts).astype(np.uint32), f.post_tf[sel].copy(), f.post_len[sel].copy())
        out.append(StringIndexData([fp], docs.shape[0], docs.shape[0], docs))
    return out


@pytest.mark.parametrize("mode,offset", [(MODE_FULLTEXT, 0), (MODE_HYBRID, 0), (MODE_HYBRID, 3)])
def test_multi_index_union(gpu_ctx, orc, mode, offset):
    n, dim, vocab, B, parts, limit = 24000, 384, 1500, 12, 3, 10
    rows = synth.make_vectors(n, dim, seed=51)
    qv, _ = synth.make_vector_queries(rows, B, seed=52)
    data = synth.make_text_corpus(n, vocab, seed=53)
    texts = synth.make_text_queries(vocab, B, seed=54)
    shards = _split_text(data, parts)
    per, union = [], [dict() for _ in range(B)]
    counts = np.zeros(B, np.int64)
    for i, sd in enumerate(shards):
        docs = sd.row_doc_ids
        emb = ob.EmbeddingFieldStorage(gpu_ctx, "BGESmall")
        emb.insert_batch(docs, rows[docs.astype(np.int64)])
        strs = ob.StringFieldStorage(gpu_ctx, sd)
        tsc = ob.TokenScoreContext(gpu_ctx, emb, strs)
        p = ob.TokenScoreParams(mode=mode, limit_hint=limit + offset, offset=0, vector_limit=limit, similarity=0.0)
        per.append(tsc.execute_batch_arrays(p, texts, qv))
        # the oracle, piece by piece, with the reference's depths: vector top-`limit`, then the whole per-index map
        ix, st = orc.StrIndex(sd), orc.EmbStore(rows[docs.astype(np.int64)], row_doc_ids=docs)
        for q in range(B):
            ft = orc.fulltext(ix, texts[q])
            m = ft if mode == MODE_FULLTEXT else orc.hybrid_combine(orc.vector(st, qv[q], limit, 0.0), ft)
            counts[q] += len(m[0])
            for d, s in zip(*m):
                if s == s:
                    union[q][int(d)] = np.float32(s)
        emb.close(); strs.close()
    hits = ob.merge_index_results(per, limit, offset)
    for q in range(B):
        exp = sorted(union[q].items(), key=lambda kv: (-kv[1], kv[0]))[offset:offset + limit]
        assert hits[q].count == int(counts[q])
        assert_topk_equal(hits[q].doc_ids, hits[q].scores, np.asarray([d for d, _ in exp], np.uint64),
                          np.asarray([s for _, s in exp], np.float32))
