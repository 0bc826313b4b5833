"""Pin the oracle's restatement of the verification arithmetic (EPnP, LM refinement, RANSAC driver,
cv::RNG) against known answers produced with the real OpenCV primitives of this image
(tests/golden/pnp_golden.json, made by tests/golden/make_pnp_golden.py)."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle_py as orc

GOLD = json.loads((Path(__file__).parent / "golden" / "pnp_golden.json").read_text())
K4 = GOLD["K"]


@pytest.mark.parametrize("case", [c for c in GOLD["primitives"] if c["kind"] == "epnp6"])
def test_epnp_minimal_sample_matches_opencv(case):
    ok, r, t = orc.solve_pnp_epnp(case["X"], case["uv"], K4)
    assert ok
    assert np.allclose(r, case["rvec"], atol=1e-9) and np.allclose(t, case["tvec"], atol=1e-9)


@pytest.mark.parametrize("case", [c for c in GOLD["primitives"] if c["kind"] == "iterative"])
def test_iterative_refinement_matches_opencv(case):
    r, t = orc.solve_pnp_iterative(case["X"], case["uv"], K4, case["rvec0"], case["tvec0"])
    assert np.allclose(r, case["rvec"], atol=1e-6) and np.allclose(t, case["tvec"], atol=1e-6)


@pytest.mark.parametrize("case", GOLD["ransac"], ids=lambda c: f"n{c['n']}_refine{c['refine']}")
def test_ransac_inliers_and_pose_match_opencv_replay(case):
    ok, r, t, inl, iters = orc.pnp_ransac(case["X"], case["uv"], K4, case["iterations"], case["reproj"], case["min_inliers"], case["refine"])
    assert ok == case["ok"]
    assert iters == case["iterations_run"]          # same RNG stream, same adaptive iteration count
    assert inl.tolist() == case["inliers"]          # bit-exact inlier set
    assert np.allclose(r, case["rvec"], atol=1e-6) and np.allclose(t, case["tvec"], atol=1e-6)


def test_sym_eigen_ql_against_jacobi_and_numpy():
    """The Householder+QL solver the restatement uses, checked against the independent cyclic-Jacobi solver kept in
    oracle/pnp_math.h and against numpy (LAPACK) on Gram matrices of the sizes the PnP path decomposes."""
    rng = np.random.default_rng(5)
    for n in (3, 4, 5, 6, 12):
        for trial in range(20):
            m = rng.standard_normal((n if trial % 2 else max(n - 2, 2), n))   # every other one rank-deficient
            a = m.T @ m * 10.0 ** rng.integers(-3, 4)
            w0, v0 = orc.sym_eigen(a, 0)
            w1, v1 = orc.sym_eigen(a, 1)
            wn = np.linalg.eigvalsh(a)[::-1]
            scale = max(abs(wn[0]), 1e-300)
            assert np.all(np.diff(w0) <= 0)
            assert np.allclose(w0, wn, atol=1e-13 * scale, rtol=0)
            assert np.allclose(w0, w1, atol=1e-13 * scale, rtol=0)
            assert np.allclose(v0 @ v0.T, np.eye(n), atol=1e-13)
            assert np.allclose(v0 @ a @ v0.T, np.diag(w0), atol=1e-12 * scale)
