"""CPU oracle of rtabmap::BayesFilter — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py CPU legs).

Restates, dense and literal, corelib/src/BayesFilter.cpp:
  computePosterior   :145-270   prior = prediction * lastPosterior; posterior = likelihood .* prior, normalised
  generatePrediction :301-420   one column per place: LC values at the place's graph neighbours (by margin), see addNeighborProb :272-299
  normalize          :437-505   missing neighbour mass onto the diagonal, uniform value for all other places, renormalisation,
                                epsilon cut, virtual-place row
  updatePosterior    :712-737   last posterior re-keyed to the ids of this call (new ids get 0, or 1 for the very first call)
with float (CV_32FC1) storage and the reference's statement order.  The incremental updatePrediction (:507-710,
Bayes/FullPredictionUpdate=false) is the reference's optimisation of the same matrix and is not restated.

Pinning.  The reference holds one vector for this class, archive/2010-LoopClosure/Tests/TestBayesFilter.m (a 10 x 10 table of
floor(1000 * posterior)).  Its prediction matrices come from the 2010 generatePrediction.m, whose pattern format (separate backward /
forward values, no renormalisation) predates the C++ of this tree, so the table pins the RECURSION (computePosterior + updatePosterior:
tests/test_oracle_bayes.py feeds it the 2010 matrices, tests/golden/bayes_golden.json) and NOT generatePrediction / normalize, which
are restated here from BayesFilter.cpp:301-505 — parity unpinned for those two (oracle-vs-CUDA only).

The map graph is the caller's (Memory::getNeighborsId, Memory.cpp; graph code is outside the hot path): `neighbors(id)` returns
{neighbour id: margin} for margins 0..len(prediction_lc)-2, the node itself at margin 0, places in the short-term memory already removed.
"""
from __future__ import annotations

import numpy as np

DEFAULT_PREDICTION_LC = [0.1, 0.36, 0.30, 0.16, 0.062, 0.0151, 0.00255, 0.000324, 2.5e-05, 1.3e-06, 4.8e-08, 1.2e-09, 1.9e-11, 2.2e-13, 1.7e-15,
                         8.5e-18, 2.9e-20, 6.9e-23]  # Parameters.h:363


class BayesFilterOracle:
    def __init__(self, prediction_lc=None, virtual_place_prior: float = 0.9):
        self.lc = [float(v) for v in (prediction_lc if prediction_lc is not None else DEFAULT_PREDICTION_LC)]
        self.vpp = np.float32(virtual_place_prior)
        f = np.float32
        total = f(0)
        eps = None
        for j, v in enumerate(self.lc):  # setPredictionLC :103-111 (float accumulation of doubles)
            total = f(total + v)
            if j == 0 or v < eps:
                eps = v
        self.total = total
        self.eps = f(eps)
        self.posterior: dict[int, np.float32] = {}

    def reset(self):
        self.posterior = {}

    # ---- generatePrediction + addNeighborProb + normalize (full update) ---------------------------------------------
    def prediction(self, ids, neighbors) -> np.ndarray:
        f = np.float32
        n = len(ids)
        P = np.zeros((n, n), np.float32)  # P[row, col]: data[col + row * cols]
        index = {int(i): k for k, i in enumerate(ids) if i > 0}
        vp_used = ids[0] < 0
        done = set()
        for i, pid in enumerate(ids):
            pid = int(pid)
            if pid in done:
                continue
            if pid > 0:
                nb = dict(sorted(neighbors(pid).items()))
                loop_margin = [k for k, m in nb.items() if m == 0 and k in index]
                assert loop_margin, f"No 0 margin neighbor for signature {pid}"
                for lid in loop_margin:
                    col = index[lid]
                    s = f(0)
                    for k, m in nb.items():
                        if k >= 0 and k in index:
                            v = f(self.lc[m + 1])
                            P[index[k], col] = v
                            s = f(s + v)
                    done.add(lid)
                    self._normalize(P, col, s, vp_used)
            else:
                if self.vpp > 0:
                    if n > 1:
                        P[0, i] = self.vpp
                        P[1:, i] = f((1.0 - float(self.vpp)) / (n - 1))
                    else:
                        P[0, i] = 1
                else:
                    P[:, i] = f(1.0 / n) if n > 1 else f(1)
        return P

    def _normalize(self, P, col, s, vp_used):
        f = np.float32
        n = P.shape[0]
        lc0 = self.lc[0]
        if s < float(self.total) - lc0:
            delta = f(float(self.total) - lc0 - float(s))
            P[col, col] = f(P[col, col] + delta)
            s = f(s + delta)
        other = f(0)
        if self.total < 1:
            other = f(f(1.0) - self.total)
        first = 1 if vp_used else 0
        if other > 0 and n > 1:
            value = f(other / f(n - 1))
            for j in range(first, n):
                if P[j, col] == 0:
                    P[j, col] = value
                    s = f(s + value)
        max_norm = f(1 - (lc0 if vp_used else 0))
        if s < float(max_norm) - 0.0001 or s > float(max_norm) + 0.0001:
            scale = f(max_norm / s)
            for j in range(first, n):
                P[j, col] = f(P[j, col] * scale)
                if P[j, col] < self.eps:
                    P[j, col] = 0
        if vp_used:
            P[0, col] = f(lc0)

    # ---- computePosterior ------------------------------------------------------------------------------------------------
    def compute_posterior(self, ids, likelihood, neighbors=None, prediction=None) -> np.ndarray:
        """`prediction` (dense, float32 [n, n]) overrides generatePrediction — used to replay the reference's 2010 test matrices."""
        ids = [int(i) for i in ids]
        assert ids == sorted(ids), "uKeys(likelihood): ascending ids"
        P = np.asarray(prediction, np.float32) if prediction is not None else self.prediction(ids, neighbors)
        first = len(self.posterior) == 0
        last = np.array([self.posterior.get(i, np.float32(1 if first else 0)) for i in ids], np.float32)  # updatePosterior
        prior = P @ last                                   # cv::Mat product, float
        post = (np.asarray(likelihood, np.float32) * prior).astype(np.float32)
        s = np.float32(0)
        for v in post:                                     # sum in id order (:236-247)
            s = np.float32(s + v)
        if s != 0:
            post = (post / s).astype(np.float32)
        self.posterior = {i: post[k] for k, i in enumerate(ids)}
        return post


def chain_neighbors(n_levels: int, present=None, loops=None):
    """Neighbour function of a trajectory graph: place k is linked to k-1 and k+1; `loops` adds loop-closure links {a: b} (margin 1
    both ways, as any other link).  BFS by margin, limited to `present` ids when given."""
    loops = loops or {}
    adj_extra: dict[int, set] = {}
    for a, b in loops.items():
        adj_extra.setdefault(a, set()).add(b)
        adj_extra.setdefault(b, set()).add(a)

    def nb(pid: int):
        out = {pid: 0}
        frontier = [pid]
        for m in range(1, n_levels + 1):
            nxt = []
            for u in frontier:
                for v in [u - 1, u + 1, *adj_extra.get(u, ())]:
                    if v >= 1 and v not in out and (present is None or v in present):
                        out[v] = m
                        nxt.append(v)
            frontier = nxt
        return out

    return nb


def prediction_columns(ids, neighbors):
    """The neighbour lists BayesFilter::generatePrediction ends up using for every column, in CSR form for lcd_bayes_compute_posterior:
    (col_ptr int64 [n+1], nbr_row int32, nbr_level int32).  Same walk as BayesFilterOracle.prediction: places are visited in id order, a
    place and its margin-0 partners (loop-closure links) share the neighbour tree of whichever is visited first (:352-391)."""
    ids = [int(i) for i in ids]
    index = {i: k for k, i in enumerate(ids) if i > 0}
    cols: dict[int, list] = {}
    done = set()
    for pid in ids:
        if pid <= 0 or pid in done:
            continue
        nb = dict(sorted(neighbors(pid).items()))
        entries = [(index[k], m) for k, m in nb.items() if k >= 0 and k in index]
        for lid in [k for k, m in nb.items() if m == 0 and k in index]:
            cols[index[lid]] = entries
            done.add(lid)
    col_ptr = [0]
    rows, levels = [], []
    for c in range(len(ids)):
        for r, m in cols.get(c, []):
            rows.append(r)
            levels.append(m)
        col_ptr.append(len(rows))
    return np.array(col_ptr, np.int64), np.array(rows, np.int32), np.array(levels, np.int32)
