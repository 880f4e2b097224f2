"""Generates tests/golden/bm25_known_answers.json.

The reference cannot be compiled or imported here (pure Rust, no cargo).  Its own numeric
pins for this path are the closed-form known-answer tests in
src/collection_manager/bm25.rs:534-563 (test_bm25f_scorer_basic, tol 1e-6) and :912-983
(test_canonical_bm25f_single_term_two_fields, tol 1e-5).  This script re-evaluates those
closed forms exactly as the Rust tests write them (f32 arithmetic, ln_1p) and records
inputs + expected values; the inequality pins (:566-909, 986-1043) are recorded as relations.
"""
import json
import os

import numpy as np

f = np.float32


def ln_1p(x):
    return f(np.log1p(f(x)))


cases = {}

# bm25.rs:534-563
ratio = f(f(f(100.0) - f(10.0)) + f(0.5)) / f(f(10.0) + f(0.5))
idf = ln_1p(ratio)
ntf = f(5.0)
expected = f(f(idf * f(f(1.2) + f(1.0))) * ntf) / f(f(1.2) + ntf)
cases["scorer_basic"] = dict(src="bm25.rs:534-563", tf=5, field_len=100, avg_len=100.0, total_docs=100.0, df=10,
                             k=1.2, weight=1.0, b=0.75, boost=1.0, expected=float(expected), tol=1e-6)

# bm25.rs:912-983
t_ntf = f(2.0) / f(f(f(1.0) - f(0.75)) + f(f(0.75) * f(f(10.0) / f(8.0))))
c_ntf = f(1.0) / f(f(f(1.0) - f(0.75)) + f(f(0.75) * f(f(200.0) / f(150.0))))
S = f(f(f(2.0) * t_ntf) + f(f(1.0) * c_ntf))
ratio = f(f(f(100.0) - f(10.0)) + f(0.5)) / f(f(10.0) + f(0.5))
idf = ln_1p(ratio)
k = f(1.2)
expected2 = f(f(idf * f(k + f(1.0))) * S) / f(k + S)
cases["canonical_two_fields"] = dict(
    src="bm25.rs:912-983", k=1.2, corpus_docs=100, term_docs=10,
    fields=[dict(weight=2.0, b=0.75, tf=2, len=10, avg=8.0), dict(weight=1.0, b=0.75, tf=1, len=200, avg=150.0)],
    expected=float(expected2), tol=1e-5, title_ntf=float(t_ntf), content_ntf=float(c_ntf), S=float(S), idf=float(idf))

# inequality pins
cases["relations"] = [
    dict(src="bm25.rs:566-618", what="boost 2.0 > 1.0 > 0.5 on the legacy add path"),
    dict(src="bm25.rs:673-716", what="b=0.2 scores higher than b=0.9 when len=2*avg"),
    dict(src="bm25.rs:867-909", what="score increases with weight 0.5<1<1.5<2<3; ratio(2x/1x) in (1,1.5)"),
    dict(src="bm25.rs:986-1043", what="canonical BM25F <= sum of per-field BM25"),
]

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bm25_known_answers.json")
json.dump(cases, open(out, "w"), indent=1)
print("wrote", out)
