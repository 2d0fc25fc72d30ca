"""CPU restatements of two host/device scheduling rules added in round 2 (tree.cu):
  * tree_enqueue(): the first level that goes behind the build graph's IF nodes -- a uniform tree must end above
    it (so the conditional bodies stay idle), a deep tree below it (so they are exercised);
  * ends_first(): the launch order of the walk's CTAs is a permutation that starts with the last sixteenth of the
    target range (the stragglers of a rank's shard) and then runs the range from its start."""
import numpy as np
import pytest

from inputs import masses_np, uniform_mt, uniform_np


def shallow_level(n, leaf_cap):                      # tree_enqueue: lv = levels n / leaf_cap leaves fill, + 3
    lv, cells = 0, 1
    while cells * leaf_cap < n and lv < 64:
        cells *= 8
        lv += 1
    return lv + 3


def ends_first(cta, n_ctas):                         # tree.cu: ends_first()
    tail = (n_ctas + 15) >> 4
    return n_ctas - tail + cta if cta < tail else cta - tail


@pytest.mark.parametrize("n", [2000, 20000, 200000])
def test_uniform_trees_end_above_the_conditional_levels(oracle, n):
    t = oracle.tree_build(uniform_mt(n, seed=3), masses_np(n, seed=4))
    assert t.depth <= shallow_level(n, 8), (t.depth, shallow_level(n, 8))      # depth = number of levels in use


def test_box_convention_tree_reaches_the_conditional_levels(oracle):
    n = 20000
    t = oracle.tree_build(uniform_np(n, seed=5, lo=0.0, hi=100.0), masses_np(n, seed=6))
    assert t.depth == 21 and t.depth > shallow_level(n, 8)


@pytest.mark.parametrize("n_ctas", [1, 2, 15, 16, 17, 255, 2048, 4097])
def test_ends_first_is_a_permutation_that_starts_with_the_tail(n_ctas):
    order = [ends_first(c, n_ctas) for c in range(n_ctas)]
    assert sorted(order) == list(range(n_ctas))
    tail = (n_ctas + 15) >> 4
    assert order[:tail] == list(range(n_ctas - tail, n_ctas))
    assert order[tail:] == list(range(n_ctas - tail))
