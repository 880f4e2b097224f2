"""Synthetic corpora of the BASELINE.json shapes, seeded and deterministic.

numpy PCG64 streams (seeded) stand in for the splitmix/xoshiro wording of §8d: the same
arrays feed the GPU path and the CPU oracle, so the generator is not part of parity.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .types import FieldPostings, StringIndexData, TextQuery

SEED_VECTORS = 0x0A11CE5EED01
SEED_VQUERIES = 0xBEEF0001
SEED_TEXT = 0x5EED7E47
SEED_TQUERIES = 0xBEEF0002


def make_vectors(n: int, dim: int, seed: int = SEED_VECTORS, chunk: int = 65536) -> np.ndarray:
    """iid N(0,1) components, each row scaled by exp(0.25*N(0,1)) (non-unit norms: the reference
    mean-pools without L2 normalisation, scripts/src/embeddings/embeddings.py:39-42)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        x = rng.standard_normal((m, dim), dtype=np.float32)
        s = np.exp(0.25 * rng.standard_normal(m, dtype=np.float32)).astype(np.float32)
        out[i:i + m] = x * s[:, None]
    return out


def make_clustered_vectors(n: int, dim: int, n_centroids: int = 2000, sigma: float = 0.1, seed: int = SEED_VECTORS,
                           chunk: int = 65536) -> np.ndarray:
    """Near-duplicate clusters (the normal case for real text embeddings, and the adversarial one for a
    low-precision candidate selection): row = c_k + sigma*N(0,I) with c_k ~ N(0,I), k uniform, so
    cos(row, c_k) ~ 1/sqrt(1+sigma^2) (0.995 at sigma = 0.1) and two members of one cluster sit at
    ~1/(1+sigma^2) = 0.990 with a spread of ~1e-3: hundreds of rows within 0.01 of every query's best
    hit.  Rows of a cluster are scattered over the row space; norms vary like make_vectors'."""
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((n_centroids, dim), dtype=np.float32)
    out = np.empty((n, dim), np.float32)
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        k = rng.integers(0, n_centroids, size=m)
        x = cent[k] + np.float32(sigma) * rng.standard_normal((m, dim), dtype=np.float32)
        s = np.exp(0.25 * rng.standard_normal(m, dtype=np.float32)).astype(np.float32)
        out[i:i + m] = x * s[:, None]
    return out


def make_vector_queries(rows: np.ndarray, b: int, seed: int = SEED_VQUERIES, noise: float = 0.3) -> Tuple[np.ndarray, np.ndarray]:
    """q = x_j + 0.3*N(0,I) for uniformly drawn j; returns (queries, j)."""
    rng = np.random.default_rng(seed)
    j = rng.integers(0, rows.shape[0], size=b)
    q = rows[j] + noise * rng.standard_normal((b, rows.shape[1]), dtype=np.float32)
    return q.astype(np.float32), j


def _zipf_cdf(vocab: int, s: float) -> np.ndarray:
    p = 1.0 / np.power(np.arange(1, vocab + 1, dtype=np.float64), s)
    c = np.cumsum(p)
    return c / c[-1]


def make_text_corpus(n_docs: int, vocab: int, seed: int = SEED_TEXT, zipf_s: float = 1.0,
                     mean_len: float = 32.0, sigma: float = 0.5, chunk_docs: int = 1 << 20) -> StringIndexData:
    """Single string field: doc length L = clamp(round(exp(N(ln mean_len, sigma^2))), 1, 65535)
    tokens (u16 like string_field.rs:162), term rank ~ Zipf(s) over `vocab`, tf = multiplicity."""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.rint(np.exp(rng.normal(np.log(mean_len), sigma, n_docs))), 1, 65535).astype(np.int64)
    cdf = _zipf_cdf(vocab, zipf_s)
    keys = []
    for d0 in range(0, n_docs, chunk_docs):
        d1 = min(n_docs, d0 + chunk_docs)
        ntok = int(lens[d0:d1].sum())
        terms = np.searchsorted(cdf, rng.random(ntok), side="right").astype(np.uint64)
        np.minimum(terms, np.uint64(vocab - 1), out=terms)
        docs = np.repeat(np.arange(d0, d1, dtype=np.uint64), lens[d0:d1])
        k = (terms << np.uint64(32)) | docs
        k.sort()
        keys.append(k)
    allk = np.concatenate(keys) if len(keys) > 1 else keys[0]
    if len(keys) > 1:
        allk.sort(kind="stable")
    uk, tf = np.unique(allk, return_counts=True)
    del allk
    post_term = (uk >> np.uint64(32)).astype(np.int64)
    post_row = (uk & np.uint64(0xffffffff)).astype(np.uint32)
    post_tf = np.minimum(tf, 65535).astype(np.uint16)
    post_len = lens[post_row].astype(np.uint16)
    offs = np.zeros(vocab + 1, np.uint64)
    offs[1:] = np.cumsum(np.bincount(post_term, minlength=vocab)).astype(np.uint64)
    field = FieldPostings(float(lens.mean()), offs, post_row, post_tf, post_len)
    return StringIndexData([field], n_docs, n_docs, None)


def make_text_queries(vocab: int, b: int, terms_per_query: int = 3, seed: int = SEED_TQUERIES,
                      zipf_s: float = 1.0) -> List[TextQuery]:
    """Each query: `terms_per_query` distinct term ids drawn from the same Zipf."""
    rng = np.random.default_rng(seed)
    cdf = _zipf_cdf(vocab, zipf_s)
    out = []
    for _ in range(b):
        ids: List[int] = []
        while len(ids) < terms_per_query:
            t = int(min(np.searchsorted(cdf, rng.random(), side="right"), vocab - 1))
            if t not in ids:
                ids.append(t)
        out.append(TextQuery.single_terms(ids))
    return out
