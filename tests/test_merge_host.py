"""oc_merge_results (host): the union of per-index result lists == one top_n over the union of the per-index
score maps (search.rs:304-338, 482-498).  Runs without a GPU."""
import numpy as np

import oramacore_b200 as ob


def test_merge_equals_sort_of_the_union():
    rng = np.random.default_rng(2)
    B, k, limit, offset = 7, 3, 5, 2
    stride = limit + offset
    per, union = [], [dict() for _ in range(B)]
    counts_total = np.zeros(B, np.uint64)
    for i in range(k):
        docs = np.zeros((B, stride), np.uint64)
        scores = np.zeros((B, stride), np.float32)
        n = rng.integers(0, stride + 1, size=B).astype(np.uint32)
        cnt = (n + rng.integers(0, 50, size=B)).astype(np.uint64)
        for q in range(B):
            s = np.sort(rng.choice([0.5, 1.0, 1.5, 2.0, 2.5], size=n[q]).astype(np.float32))[::-1]   # many ties
            d = (rng.choice(1000, size=n[q], replace=False) * k + i).astype(np.uint64)                 # disjoint per index
            order = np.lexsort((d, -s))
            docs[q, :n[q]], scores[q, :n[q]] = d[order], s[order]
            for dd, ss in zip(d, s):
                union[q][int(dd)] = float(ss)
        counts_total += cnt
        per.append((docs, scores, n, cnt))
    hits = ob.merge_index_results(per, limit, offset)
    for q in range(B):
        exp = sorted(union[q].items(), key=lambda kv: (-kv[1], kv[0]))[offset:offset + limit]
        assert hits[q].doc_ids.tolist() == [d for d, _ in exp]
        assert hits[q].scores.tolist() == [s for _, s in exp]
        assert hits[q].count == int(counts_total[q])
