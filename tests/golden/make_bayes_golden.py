"""Golden vector of the Bayes filter recursion: the reference's own archive/2010-LoopClosure/Tests/TestBayesFilter.m.

The .m file asserts a 10 x 10 table floor(1000 * posterior) produced by
    prior = likelihood .* (generatePrediction(0.9, LC, m)' * [prior; 0]);  prior = prior / sum(prior)
with likelihood = ones and the 2010 pattern LC = [0.1 0.24 0.18 0.18 0.1 0.1 0.04 0.04 0.01 0.01].  This script restates
archive/2010-LoopClosure/Bayes/generatePrediction.m (the 2010 pattern format: [virtual place, loop closure, n-1, n+1, n-2, n+2, ...]) to
produce the ten prediction matrices, copies the asserted table from the .m file, and checks in float64 that the recursion reproduces it.

    python tests/golden/make_bayes_golden.py   ->  tests/golden/bayes_golden.json
"""
import json
import re
from pathlib import Path

import numpy as np

REF = Path("/root/reference/archive/2010-LoopClosure/Tests/TestBayesFilter.m")
OUT = Path(__file__).resolve().parent / "bayes_golden.json"


def generate_prediction_2010(new_place, lc, m):
    P = np.zeros((m + 1, m + 1))
    P[0, :] = [new_place] + [(1 - new_place) / m] * m if m > 0 else [new_place]
    for i in range(2, m + 2):
        y = np.zeros(m + 1)
        y[0] = lc[0]
        y[1:] = (1 - sum(lc)) / m
        y[i - 1] += lc[1]
        added = lc[1]
        n = i
        for k in range(3, len(lc) + 1, 2):
            n -= 1
            if n > 1:
                y[n - 1] += lc[k - 1]
                added += lc[k - 1]
            else:
                break
        n = i
        for k in range(4, len(lc) + 1, 2):
            n += 1
            if n <= len(y):
                y[n - 1] += lc[k - 1]
                added += lc[k - 1]
            else:
                break
        total = sum(lc[1:])
        if added < total:
            y[i - 1] += total - added
        P[i - 1, :] = y
    return P


def main():
    src = REF.read_text()
    lc = [float(v) for v in re.search(r"predictionLC = \[([^\]]+)\]", src).group(1).split()]
    np_prior = float(re.search(r"predictionNP = ([0-9.]+)", src).group(1))
    rows = re.search(r"computePosteriorResult - \[([^\]]+)\]", src).group(1).strip().strip(";").split(";")
    table = [[int(v) for v in r.split(",")] for r in rows]
    n_iter = len(table)
    preds = []
    prior = None
    for i in range(1, n_iter + 1):
        P = generate_prediction_2010(np_prior, lc, i - 1)
        preds.append(P.T.tolist())  # the C++ layout: column = last state
        prior = np.array([1.0]) if i == 1 else np.concatenate([prior, [0.0]])
        prior = np.ones(i) * (P.T @ prior)
        prior /= prior.sum()
        x = np.concatenate([prior * 1000, np.zeros(n_iter - i)])
        t = np.array(table[i - 1], float)
        # floor() of values that sit on an integer (0.1 * 1000) depends on the last bit of MATLAB's doubles: compare as intervals
        assert np.all((x > t - 1e-6) & (x < t + 1 + 1e-6)), (i, x, t)
    OUT.write_text(json.dumps({"source": "archive/2010-LoopClosure/Tests/TestBayesFilter.m", "prediction_lc_2010": lc, "virtual_place_prior": np_prior,
                               "table_floor_1000": table, "predictions": preds}))
    print(f"{OUT}: {n_iter} iterations reproduced")


if __name__ == "__main__":
    main()
