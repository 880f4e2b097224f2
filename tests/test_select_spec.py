"""Specification of block_select_largest (oramacore_b200/csrc/oc_common.cuh) restated in Python and
checked against a sort: radix select on unique non-zero 64-bit keys, most significant DIFFERING
byte first, stopping when the boundary bin is taken whole.  The CUDA routine is the same algorithm
with one histogram per pass in shared memory; its results are covered on the GPU by the merge /
fusion parity tests — this test pins the algorithm (termination, tie-free selection, zero keys)."""
import numpy as np
import pytest


def select_largest(keys, keep):
    keys = [int(k) for k in keys]
    nz = [k for k in keys if k]
    if len(nz) <= keep:
        return sorted(nz, reverse=True)
    o, a = 0, (1 << 64) - 1
    for k in nz:
        o |= k
        a &= k
    diff = o ^ a
    shift = ((diff | 1).bit_length() - 1) // 8 * 8
    prefix = 0 if shift == 56 else (a >> (shift + 8)) << (shift + 8)
    need, passes = keep, 0
    while shift >= 0:
        passes += 1
        hist = [0] * 256
        for k in nz:
            if shift == 56 or (k >> (shift + 8)) == (prefix >> (shift + 8)):
                hist[(k >> shift) & 255] += 1
        cum, done = 0, False
        for b in range(255, -1, -1):
            if cum + hist[b] >= need:
                prefix |= b << shift
                done = hist[b] == need - cum
                need -= cum
                break
            cum += hist[b]
        if done:
            break
        shift -= 8
    shift = max(shift, 0)
    out = [k for k in nz if (k >> shift) >= (prefix >> shift)]
    assert len(out) == keep, (len(out), keep)
    assert passes <= 8
    return sorted(out, reverse=True)


def _keys_from_scores(scores, rng):
    u = np.asarray(scores, np.float32).view(np.uint32).astype(np.uint64)
    o = np.where(u & 0x80000000, ~u & 0xFFFFFFFF, u | 0x80000000)          # f32_ordered
    idx = rng.permutation(len(scores)).astype(np.uint64)
    return (o << np.uint64(32)) | (~idx & np.uint64(0xFFFFFFFF))           # make_key: unique through the row index


@pytest.mark.parametrize("n,keep", [(300, 32), (1500, 48), (2200, 64), (4096, 64), (70, 64), (5, 10)])
def test_select_equals_sort(n, keep):
    rng = np.random.default_rng(n + keep)
    for scores in (rng.normal(0.17, 0.01, n), rng.normal(0, 1, n), np.full(n, 0.25), np.round(rng.normal(0, 1, n), 1)):
        keys = _keys_from_scores(scores, rng)
        keys[rng.integers(0, n, size=max(1, n // 10))] = 0                   # KEY_NONE holes (dropped candidates)
        exp = sorted((int(k) for k in keys if k), reverse=True)[:keep]
        assert select_largest(keys, keep) == exp


def test_select_resolves_in_the_index_bytes_when_scores_are_equal():
    rng = np.random.default_rng(3)
    keys = _keys_from_scores(np.full(1000, 0.5), rng)                      # identical scores: duplicates of one vector
    exp = sorted((int(k) for k in keys), reverse=True)[:32]
    assert select_largest(keys, 32) == exp
