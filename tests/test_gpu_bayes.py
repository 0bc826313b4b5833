"""lcd_bayes_compute_posterior (sparse prediction on the GPU) against the dense float restatement of BayesFilter.cpp
(oracle/bayes_py.py, whose recursion is pinned to the reference's TestBayesFilter.m table in tests/test_oracle_bayes.py).
Tolerance: 1e-4 relative (north_star, float work); observed ~1e-6."""
import numpy as np
import pytest

from oracle.bayes_py import DEFAULT_PREDICTION_LC, BayesFilterOracle, chain_neighbors, prediction_columns
from rtabmap_b200 import Engine
from rtabmap_b200.capi import LcdError

pytestmark = pytest.mark.gpu
LEVELS = len(DEFAULT_PREDICTION_LC) - 2


def run_stream(lc, vpp, use_vp, n0, steps, loops, rng, drop_every=0):
    eng = Engine()
    bf = BayesFilterOracle(lc, vpp)
    places = list(range(1, n0 + 1))
    worst = 0.0
    for t in range(steps):
        places.append(places[-1] + 1)                       # the map grows by one place per step ...
        if drop_every and t % drop_every == drop_every - 1:  # ... and sometimes forgets some (memory management / reactivation)
            for victim in rng.choice(places[5:-5], 3, replace=False):
                places.remove(int(victim))
        ids = ([-1] if use_vp else []) + places
        nb = chain_neighbors(len(lc) - 2, present=set(places), loops=loops)
        like = np.ones(len(ids), np.float32)
        peak = 10 + 2 * t
        like[peak] = 25.0
        like[peak - 1] = like[peak + 1] = 6.0
        like += rng.random(len(ids)).astype(np.float32) * 0.2
        col_ptr, rows, levels = prediction_columns(ids, nb)
        g = eng.bayes_compute_posterior(ids, like, col_ptr, rows, levels, lc, vpp)
        o = bf.compute_posterior(ids, like, nb)
        assert abs(float(g.sum()) - 1.0) < 1e-4
        assert np.allclose(g, o, rtol=1e-4, atol=1e-9), f"step {t}: max rel {np.max(np.abs(g - o) / np.maximum(o, 1e-12))}"
        assert int(np.argmax(g)) == int(np.argmax(o))
        worst = max(worst, float(np.max(np.abs(g - o) / np.maximum(np.abs(o), 1e-9))))
    return worst


def test_posterior_stream_with_virtual_place_and_loop_closures():
    rng = np.random.default_rng(1)
    worst = run_stream(DEFAULT_PREDICTION_LC, 0.9, True, 120, 14, {5: 90, 33: 70}, rng, drop_every=4)
    assert worst < 1e-4


def test_posterior_stream_without_virtual_place():
    rng = np.random.default_rng(2)
    run_stream(DEFAULT_PREDICTION_LC, 0.9, False, 60, 8, {}, rng)


def test_posterior_with_a_pattern_that_sums_to_one_and_zero_prior():
    """sum(PredictionLC) == 1: no uniform value for the other places; VirtualPlacePriorThr = 0: the uniform virtual column (:404-417)."""
    rng = np.random.default_rng(3)
    run_stream([0.1, 0.5, 0.2, 0.1, 0.1], 0.9, True, 40, 6, {3: 30}, rng)
    run_stream(DEFAULT_PREDICTION_LC, 0.0, True, 40, 6, {}, rng)


def test_large_map_is_sparse_work():
    """20 000 places: the reference's dense matrix would be 1.6 GB; the sparse form is a few hundred thousand entries."""
    eng = Engine()
    n = 20000
    ids = [-1] + list(range(1, n + 1))
    nb = chain_neighbors(LEVELS, present=None)
    col_ptr = [0]
    rows, levels = [], []
    for c, pid in enumerate(ids):
        if pid > 0:
            for k, m in sorted(nb(pid).items()):
                if 1 <= k <= n:
                    rows.append(k)       # position in ids == id (virtual place at 0)
                    levels.append(m)
        col_ptr.append(len(rows))
    like = np.ones(n + 1, np.float32)
    like[777] = 40.0
    post = None
    for _ in range(8):
        post = eng.bayes_compute_posterior(ids, like, col_ptr, rows, levels, DEFAULT_PREDICTION_LC, 0.9)
    assert abs(float(post.sum()) - 1.0) < 1e-4 and int(np.argmax(post)) == 777 and post[777] > 0.5


def test_rejects_malformed_input():
    eng = Engine()
    lc = DEFAULT_PREDICTION_LC
    with pytest.raises(LcdError):   # ids not ascending
        eng.bayes_compute_posterior([3, 2], [1, 1], [0, 1, 2], [0, 1], [0, 0], lc)
    with pytest.raises(LcdError):   # a column that does not list its own place
        eng.bayes_compute_posterior([1, 2], [1, 1], [0, 1, 2], [1, 1], [1, 0], lc)
    with pytest.raises(LcdError):   # margin beyond the pattern
        eng.bayes_compute_posterior([1, 2], [1, 1], [0, 1, 2], [0, 1], [0, 99], lc)
