"""The tensor-core sweeps only SELECT candidates; the merge kernel proves that the exact top-`limit`
is inside them using a rigorous bound eps on |approx - exact| of the cosine (emb_gemm.cuh:
GEMM_EPS_*).  These CPU tests pin the constants: they emulate the operand roundings in numpy
(tf32 = fp32 with the low 13 mantissa bits dropped; bf16 = round to nearest even, 8 significant
bits), check the analytic bounds on random and on adversarial vectors (which nearly attain them),
and check that the constants compiled into the library cover bound + fp32 accumulation."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACC = 1024 * 2.0 ** -23          # <= 1024 fp32 adds, truncating accumulate, relative to sum |x_i q_i| <= |x||q|


def tf32_trunc(a):
    return (np.ascontiguousarray(a, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16_rn(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return (r.astype(np.uint32) << np.uint32(16)).view(np.float32)


def cos_err(x, q, fx, fq):
    x64, q64 = x.astype(np.float64), q.astype(np.float64)
    exact = (x64 * q64).sum(-1)
    approx = (fx(x).astype(np.float64) * fq(q).astype(np.float64)).sum(-1)
    return np.abs(approx - exact) / (np.linalg.norm(x64, axis=-1) * np.linalg.norm(q64, axis=-1))


ident = lambda a: a
MODES = {   # name -> (row rounding, query rounding, analytic bound without accumulation)
    "TF32": (tf32_trunc, tf32_trunc, 2 * 2.0 ** -10 + 2.0 ** -20),
    "BF16_Q": (ident, bf16_rn, 2.0 ** -8),
    "BF16X2": (bf16_rn, bf16_rn, 2 * 2.0 ** -8 + 2.0 ** -16),
}


def _constants():
    src = open(os.path.join(ROOT, "oramacore_b200", "csrc", "emb_gemm.cuh")).read()
    return {k: float(v) for k, v in re.findall(r"constexpr float GEMM_EPS_(\w+) = ([0-9.e+-]+)f;", src)}


@pytest.mark.parametrize("mode", sorted(MODES))
def test_compiled_eps_covers_the_rigorous_bound(mode):
    c = _constants()
    assert mode in c, c
    bound = MODES[mode][2] + ACC
    assert bound <= c[mode] <= bound * 1.06, (mode, bound, c[mode])     # sound, and not needlessly loose


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("dim", [384, 768, 1024])
def test_bound_holds_on_random_vectors(mode, dim):
    fx, fq, bound = MODES[mode]
    rng = np.random.default_rng(dim)
    x = rng.standard_normal((4000, dim)).astype(np.float32)
    q = rng.standard_normal((1, dim)).astype(np.float32)
    e = cos_err(x, q, fx, fq)
    assert e.max() <= bound
    # typical errors are far below the worst case (why the proof almost always succeeds)
    assert np.median(e) < bound / 10


@pytest.mark.parametrize("mode", sorted(MODES))
def test_bound_is_nearly_attained(mode):
    """Adversarial rows: every element sits just below the next representable value (tf32) / just
    below the rounding midpoint (bf16), all errors aligned: the analytic bound is tight, a smaller
    constant would make the proof unsound."""
    fx, fq, bound = MODES[mode]
    dim = 768
    if mode == "TF32":
        v = np.full(dim, np.uint32(0x3F801FFF)).view(np.float32)             # 1 + (2^13 - 1) 2^-23
    else:
        v = np.full(dim, np.float32(1.0 + 2.0 ** -8 - 2.0 ** -20), np.float32)   # rounds down to 1.0
    e = cos_err(v[None, :], v[None, :], fx, fq)[0]
    assert bound * 0.97 <= e <= bound, (mode, e, bound)


def test_candidate_depth_choice():
    """P[proof fails] per query for random-like data = P[cos_(limit) - cos_(K') < eps] (order statistics of
    n N(0, 1/d) cosines; Renyi representation of the top order statistics).  K' = 32 is ample for tf32;
    the bf16 arithmetics need K' = 64 (capi.cu: keep)."""
    from scipy.stats import norm
    rng = np.random.default_rng(0)

    def p_unproven(n, d, K, eps, limit=10, trials=100000):
        S = np.cumsum(rng.exponential(size=(trials, K)), axis=1)
        x = norm.isf(S / n) / np.sqrt(d)
        return float(((x[:, limit - 1] - x[:, K - 1]) < eps).mean())

    c = _constants()
    assert p_unproven(1e6, 768, 32, c["TF32"]) < 1e-4
    assert p_unproven(1e6, 768, 48, c["BF16X2"]) > 1e-3          # why 48 is not enough for the converting sweep
    assert p_unproven(1e6, 768, 64, c["BF16X2"]) < 1e-4
    assert p_unproven(1e7, 1024, 32, c["BF16_Q"]) > 1e-3         # bf16 store, BASELINE configs[4] shape
    assert p_unproven(1e7, 1024, 64, c["BF16_Q"]) < 1e-4
