"""bench.py's fp64 recall checkers against the C oracle on small corpora (CPU)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oramacore_b200 import synth


def test_fp64_vector_topk_matches_oracle_f64(orc):
    rows = synth.make_clustered_vectors(3000, 96, n_centroids=20, seed=3)
    qv, _ = synth.make_vector_queries(rows, 9, seed=4)
    vi, vs = bench.fp64_vector_topk(rows, qv, 10, chunk=700)
    st = orc.EmbStore(rows)
    for i in range(9):
        ed, ec = orc.vector_f64(st, qv[i], 10)
        assert np.allclose(vs[i], ec, atol=1e-12)
        assert set(vi[i].tolist()) == set(int(x) for x in ed)


def test_fp64_hybrid_topk_close_to_fp32_oracle(orc):
    n, dim, vocab = 4000, 64, 300
    rows = synth.make_vectors(n, dim, seed=5)
    qv, _ = synth.make_vector_queries(rows, 12, seed=6)
    data = synth.make_text_corpus(n, vocab, seed=7)
    texts = synth.make_text_queries(vocab, 12, seed=8)
    vi, vs = bench.fp64_vector_topk(rows, qv, 10)
    ix, st = orc.StrIndex(data), orc.EmbStore(rows)
    sb = orc.SearchBatch(ix, st)
    for i in range(12):
        sb.add(2, limit=10, similarity=-1.0, q_vec=qv[i], text=texts[i])
    od, os_, on, oc = sb.run(2)
    for i in range(12):
        ed, es = bench.fp64_hybrid_topk(data, texts[i], vi[i], vs[i], 10)
        assert np.allclose(es, os_[i, :on[i]], atol=2e-5), (es, os_[i])
        hit, tot = bench.recall_hits(od[i, :on[i]], ed, es, os_[i, :on[i]])
        assert hit == tot
