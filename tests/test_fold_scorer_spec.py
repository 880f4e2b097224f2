"""Specification of the register-folded BM25 scorers (bm25_tile3_kernel / bm25_warp_kernel, oramacore_b200/csrc/bm25.cuh)
restated in numpy and checked against the accumulator model (what bm25_tile2_kernel and the reference's per-document
`score += contribution` in token order compute).  CPU only: the CUDA kernels are compared with the oracle and with K3b
under -m gpu (tests/test_gpu_tile3.py); this file pins the ALGORITHM:

  * ownership: a row that appears in list tokens is scored by the posting of the FIRST list token holding it; every
    other row only receives dense contributions; no row is scored twice or missed;
  * fold order: contributions are added in token order with one rounding per add, an absent dense entry adds +0.0
    (no bit changes), a skipped (NaN) contribution adds nothing -> bit-identical fp32 sums;
  * threshold validity: the warm-start seed (n_keep-th best key of ANY set of scored rows, minus one) and the overflow
    redo (n_keep-th best of the first `cap` arrivals, minus one) never prune a member of the true top n_keep."""
import numpy as np
import pytest

TILE = 8192


def f32_ordered(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return np.where(u & 0x80000000, ~u & 0xFFFFFFFF, u | 0x80000000)


def make_key(score, row):
    return (f32_ordered(score) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(row, np.uint64))


def make_item(rng, n_tok, p_dense=0.4, filtered=0.0):
    """Tokens of one (tile, query) item: dense float32 arrays (0.0 = absent) or posting lists (rows ascending, payload c,
    NaN = skipped contribution)."""
    ok = rng.random(TILE) >= filtered
    toks = []
    for _ in range(n_tok):
        if rng.random() < p_dense:
            d = np.zeros(TILE, np.float32)
            rows = np.flatnonzero(rng.random(TILE) < rng.uniform(0.07, 0.9))
            d[rows] = rng.uniform(0.01, 3.0, len(rows)).astype(np.float32)
            d[~ok] = 0.0                                           # the precompute bakes the row check into dense arrays
            toks.append(("dense", d))
        else:
            n = int(rng.integers(0, 600))
            rows = np.sort(rng.choice(TILE, size=n, replace=False)).astype(np.uint32)
            c = rng.uniform(0.01, 9.0, n).astype(np.float32)
            c[rng.random(n) < 0.02] = np.nan                       # is_normal(ntf) failed: contribution skipped
            toks.append(("list", rows, c))
    return toks, ok


def accumulate(toks, ok):
    """The accumulator model: score[row] += c, token by token (K3b / the reference's hash map)."""
    s = np.zeros(TILE, np.float32)
    for t in toks:
        if t[0] == "dense":
            s = (s + t[1]).astype(np.float32)
        else:
            _, rows, c = t
            keep = ok[rows] & ~np.isnan(c)
            s[rows[keep]] = (s[rows[keep]] + c[keep]).astype(np.float32)
    return s


def fold(toks, ok):
    """K3c/K3d: per-token bitmaps + union, dense-only scan in registers, list rows folded by their first list token."""
    bm = []
    touched = np.zeros(TILE, bool)
    for t in toks:
        b = np.zeros(TILE, bool)
        if t[0] == "list":
            rows = t[1][ok[t[1]]]
            b[rows] = True
            touched |= b
        bm.append(b)
    s = np.zeros(TILE, np.float32)
    scored = np.zeros(TILE, np.int32)
    # scan: rows outside every list
    dense = [t[1] for t in toks if t[0] == "dense"]
    if dense:
        acc = dense[0].copy()                                      # 0.0 + c == c
        for d in dense[1:]:
            acc = (acc + d).astype(np.float32)
        m = ~touched
        s[m] = acc[m]
        scored[m] += 1
    else:
        scored[~touched] += 1                                      # nothing to add: stays 0 (not matched)
    # second walk: the first list token holding the row folds all tokens in order
    for j, t in enumerate(toks):
        if t[0] != "list":
            continue
        for r, cj in zip(t[1], t[2]):
            if not bm[j][r]:
                continue                                           # failed the row check
            if any(bm[jj][r] for jj in range(j)):
                continue                                           # an earlier list token owns the row
            v = np.float32(0.0)
            for i, ti in enumerate(toks):
                if ti[0] == "dense":
                    v = np.float32(v + ti[1][r])
                else:
                    if i == j:
                        ci = cj
                    elif bm[i][r]:
                        k = np.searchsorted(ti[1], r)
                        ci = ti[2][k]
                    else:
                        continue
                    if not np.isnan(ci):
                        v = np.float32(v + ci)
            s[r] = v
            scored[r] += 1
    return s, scored


@pytest.mark.parametrize("n_tok,p_dense,filtered", [(3, 0.4, 0.0), (4, 0.5, 0.3), (2, 0.0, 0.0), (3, 1.0, 0.2), (4, 0.25, 0.0), (1, 0.5, 0.0)])
def test_fold_equals_accumulate_bitwise(n_tok, p_dense, filtered):
    rng = np.random.default_rng(1000 * n_tok + int(10 * p_dense))
    for _ in range(6):
        toks, ok = make_item(rng, n_tok, p_dense, filtered)
        a = accumulate(toks, ok)
        f, scored = fold(toks, ok)
        assert np.all(scored == 1)                                 # every row scored exactly once
        assert np.array_equal(a.view(np.uint32), f.view(np.uint32))


def topn(keys, n):
    return sorted((int(k) for k in keys), reverse=True)[:n]


@pytest.mark.parametrize("n_keep", [1, 10, 32])
def test_seed_threshold_never_prunes_the_top(n_keep):
    rng = np.random.default_rng(n_keep)
    for trial in range(20):
        n = int(rng.integers(n_keep, 5000))
        scores = np.round(rng.gamma(2.0, 1.0, n), 2).astype(np.float32)   # many ties: the row index breaks them
        keys = make_key(scores, rng.permutation(n))
        sample = rng.choice(n, size=int(rng.integers(n_keep, min(n, 256) + 1)), replace=False)
        kth = topn(keys[sample], n_keep)[-1]
        tau = kth - 1                                              # "> tau" keeps the kth row itself
        kept = [int(k) for k in keys if int(k) > tau]
        assert topn(kept, n_keep) == topn(keys, n_keep)


@pytest.mark.parametrize("cap,n_keep", [(256, 10), (256, 32), (2048, 50)])
def test_overflow_redo_converges_to_the_exact_top(cap, n_keep):
    rng = np.random.default_rng(cap + n_keep)
    for order in ("random", "ascending", "descending"):
        n = 6000
        scores = rng.gamma(2.0, 1.0, n).astype(np.float32)
        keys = np.array([int(k) for k in make_key(scores, np.arange(n))], dtype=object)
        arrival = {"random": rng.permutation(n), "ascending": np.argsort(scores), "descending": np.argsort(-scores)}[order]
        tau, passes = 0, 0
        while True:
            passes += 1
            buf = [keys[i] for i in arrival if keys[i] > tau]      # candidates of this pass, in arrival order
            if len(buf) <= cap:
                break
            kth = topn(buf[:cap], n_keep)[-1] - 1                   # n_keep-th best of the first `cap` arrivals
            assert kth > tau                                       # strict progress: the loop terminates
            tau = kth
            assert passes < 64
        assert topn(buf, n_keep) == topn(keys, n_keep)
