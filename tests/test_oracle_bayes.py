"""The Bayes filter oracle (oracle/bayes_py.py) against the reference's only vector for this class,
archive/2010-LoopClosure/Tests/TestBayesFilter.m (tests/golden/bayes_golden.json, made by make_bayes_golden.py), plus structural
properties of the restated generatePrediction / normalize (columns are probability distributions, the virtual-place row, epsilon cut)."""
import json
from pathlib import Path

import numpy as np

from oracle.bayes_py import DEFAULT_PREDICTION_LC, BayesFilterOracle, chain_neighbors

GOLD = json.loads((Path(__file__).parent / "golden" / "bayes_golden.json").read_text())


def matches_table(post, row):
    """floor(1000 * posterior) == row, tolerant of values that sit exactly on an integer (0.1 * 1000 in MATLAB doubles)."""
    x = np.concatenate([np.asarray(post, np.float64) * 1000, np.zeros(len(row) - len(post))])
    t = np.asarray(row, np.float64)
    return bool(np.all((x > t - 1e-3) & (x < t + 1 + 1e-3)))


def test_recursion_reproduces_the_reference_table():
    bf = BayesFilterOracle(virtual_place_prior=GOLD["virtual_place_prior"])
    n = len(GOLD["table_floor_1000"])
    for i in range(1, n + 1):
        ids = [-1] + list(range(1, i))
        post = bf.compute_posterior(ids, np.ones(i, np.float32), prediction=np.asarray(GOLD["predictions"][i - 1]))
        assert matches_table(post, GOLD["table_floor_1000"][i - 1]), f"iteration {i}"


def test_prediction_columns_are_distributions_with_the_virtual_place_row():
    bf = BayesFilterOracle()
    ids = [-1] + list(range(1, 60))
    nb = chain_neighbors(len(DEFAULT_PREDICTION_LC) - 2, present=set(ids), loops={5: 40})
    P = bf.prediction(ids, nb)
    assert np.allclose(P.sum(0), 1.0, atol=2e-3)
    assert np.all(P[0, 1:] == np.float32(DEFAULT_PREDICTION_LC[0])) and P[0, 0] == np.float32(0.9)
    assert P[40, 5] == P[4, 5] and P[40, 5] > P[38, 5]   # the loop-closure link makes 40 a level-1 neighbour of 5


def test_posterior_follows_a_moving_likelihood_peak():
    bf = BayesFilterOracle()
    n = 80
    ids = [-1] + list(range(1, n + 1))
    nb = chain_neighbors(len(DEFAULT_PREDICTION_LC) - 2, present=set(ids))
    for t in range(6):
        like = np.ones(n + 1, np.float32)
        like[20 + t] = 30.0
        like[19 + t] = like[21 + t] = 8.0
        post = bf.compute_posterior(ids, like, nb)
        assert abs(float(post.sum()) - 1.0) < 1e-4
    assert int(np.argmax(post)) == 25 and post[25] > 0.5
