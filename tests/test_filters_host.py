"""FilterResult tree -> DocumentId bitmap (oramacore_b200/filters.py) against Python sets, and the
execute_filter rules of filter.rs:344-392 (deleted documents folded in as AND NOT)."""
import numpy as np

from oramacore_b200 import filters as F


def _as_set(bits, n):
    return {d for d in range(n) if F.contains(bits, d)}


def test_tree_matches_set_algebra():
    rng = np.random.default_rng(1)
    n = 1000
    a, b, c = (set(rng.choice(n, size=k, replace=False).tolist()) for k in (300, 450, 40))
    expr = F.Or(F.And(F.Ids(sorted(a)), F.Not(F.Ids(sorted(b)))), F.Ids(sorted(c)))
    assert _as_set(F.to_bitmap(expr, n), n) == ((a - b) | c)
    assert _as_set(F.to_bitmap(F.Not(F.Ids([])), n), n) == set(range(n))
    bits = F.to_bitmap(F.Not(F.Ids([0, 5])), 70)
    assert bits.shape == (2,) and int(bits[1]) >> 6 == 0          # padding bits beyond n_bits stay clear
    assert _as_set(F.to_bitmap(F.Ids([3, 999, 5000]), n), n) == {3, 999}   # ids outside the collection are ignored


def test_execute_filter_rules():
    n = 200
    assert F.execute_filter(None, [], n) is None                                        # nothing to filter
    assert _as_set(F.execute_filter(None, [7, 9], n), n) == set(range(n)) - {7, 9}      # NOT(deleted)
    w = F.Ids(range(0, n, 2))
    assert _as_set(F.execute_filter(w, [], n), n) == set(range(0, n, 2))
    assert _as_set(F.execute_filter(w, [4, 5], n), n) == set(range(0, n, 2)) - {4}      # AND(filter, NOT(deleted))


def test_bitmap_layout_is_what_the_oracle_and_the_kernels_read(orc):
    ids = [0, 63, 64, 129, 777]
    assert np.array_equal(F.to_bitmap(F.Ids(ids), 1000), orc.make_filter_bits(ids, 1000))
