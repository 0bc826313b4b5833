"""Make tests/golden/pnp_golden.json: known answers for the verification stage computed with the
REAL OpenCV primitives of this image (opencv-python 4.13: cv2.solvePnP EPNP / ITERATIVE,
cv2.projectPoints, cv2.Rodrigues) driven by a line-by-line Python replay of the reference's
vendored RANSAC (corelib/src/opencv/solvepnp.cpp:112-417) and refinement loop
(corelib/src/util3d_motion_estimation.cpp:810-990).  cv2.solvePnPRansac itself is a different
(USAC-era) algorithm and is NOT used.

The oracle (oracle/oracle_verify.cpp) restates those OpenCV primitives in C++; this fixture pins it.
Run here: python tests/golden/make_pnp_golden.py
"""
import json
import math
from pathlib import Path

import cv2
import numpy as np

OUT = Path(__file__).resolve().parent / "pnp_golden.json"
K = np.array([[525.0, 0, 320.0], [0, 525.0, 240.0], [0, 0, 1.0]])


class CvRng:  # cv::RNG, state (uint64)-1 (solvepnp.cpp:334)
    def __init__(self):
        self.state = 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def update_num_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, 2.2250738585072014e-308)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < 2.2250738585072014e-308:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def find_inliers(op, ip, rvec, tvec, thr):
    proj, _ = cv2.projectPoints(op.astype(np.float64), rvec, tvec, K, None)
    proj = proj.reshape(-1, 2).astype(np.float32)
    d = ip - proj
    err = np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2).astype(np.float32)
    return err <= np.float32(thr * thr)


def cv3_solve_pnp_ransac(op, ip, iterations, reproj, confidence=0.99):
    n = len(op)
    rng = CvRng()
    niters = max(iterations, 1)
    best_mask, best = None, None
    max_good = 0
    it = 0
    while it < niters:
        idx = []
        while len(idx) < 6:
            while True:
                v = rng.uniform(0, n)
                if v not in idx:
                    break
            idx.append(v)
        ok, rvec, tvec = cv2.solvePnP(op[idx], ip[idx], K, None, flags=cv2.SOLVEPNP_EPNP)
        it += 1
        if not ok:
            continue
        mask = find_inliers(op, ip, rvec, tvec, reproj)
        good = int(mask.sum())
        if good > max(max_good, 5):
            best_mask, best, max_good = mask, (rvec.copy(), tvec.copy()), good
            niters = update_num_iters(confidence, (n - good) / n, 6, niters)
    if max_good <= 0:
        return False, None, None, [], it
    return True, best[0], best[1], np.nonzero(best_mask)[0].tolist(), it


def util3d_solve_pnp_ransac(op, ip, iterations, reproj, min_inliers, refine_iterations, refine_sigma=3.0):
    ok, rvec, tvec, inliers, iters = cv3_solve_pnp_ransac(op, ip, iterations, reproj)
    if not ok:
        return False, None, None, [], iters
    min_inliers = max(min_inliers, 4)
    if len(inliers) >= min_inliers and refine_iterations > 0:
        error_threshold = np.float32(reproj)
        refine_it = 0
        prev, new = list(inliers), []
        sizes = []
        mr, mt = rvec.copy(), tvec.copy()
        changed = False
        while True:  # do { ... } while (inlier_changed && ++refine_iterations < refineIterations)
            leave = False
            _, mr, mt = cv2.solvePnP(op[prev], ip[prev], K, None, mr, mt, True, cv2.SOLVEPNP_ITERATIVE)
            sizes.append(len(prev))
            proj, _ = cv2.projectPoints(op.astype(np.float64), mr, mt, K, None)
            proj = proj.reshape(-1, 2).astype(np.float32)
            d = ip - proj
            e = np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2).astype(np.float32)
            new = np.nonzero(e <= error_threshold)[0].tolist()
            err = e[new]
            if len(new) < min_inliers:
                refine_it += 1
                if refine_it >= refine_iterations:
                    leave = True
                # else: C++ `continue` in a do-while -> the loop condition below
            else:
                m = np.float32(0)
                for v in err:
                    m = np.float32(m + v)
                m = np.float32(m / np.float32(len(err)))
                var = np.float32(0)
                if len(err) > 1:
                    s = np.float32(0)
                    for v in err:
                        s = np.float32(s + np.float32((v - m) * (v - m)))
                    var = np.float32(s / np.float32(len(err) - 1))
                error_threshold = min(np.float32(reproj), np.float32(np.float32(refine_sigma) * np.float32(math.sqrt(var))))
                changed = False
                prev, new = new, prev
                if len(new) != len(prev):
                    if len(sizes) >= min_inliers and sizes[-1] == sizes[-3] and sizes[-2] == sizes[-4]:
                        leave = True
                    else:
                        changed = True
                else:
                    changed = any(a != b for a, b in zip(prev, new))
            if leave:
                break
            refine_it += 1 if changed else 0
            if not (changed and refine_it < refine_iterations):
                break
        inliers = new
        rvec, tvec = mr, mt
    return True, rvec, tvec, inliers, iters


def make_case(rng, n, outlier_frac, noise):
    X = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(0.5, 5, n)], 1).astype(np.float32)
    rv = rng.normal(0, 0.08, 3)
    tv = rng.normal(0, 0.15, 3)
    uv, _ = cv2.projectPoints(X.astype(np.float64), rv, tv, K, None)
    uv = (uv.reshape(-1, 2) + rng.normal(0, noise, (n, 2))).astype(np.float32)
    n_out = int(outlier_frac * n)
    out_idx = rng.permutation(n)[:n_out]
    uv[out_idx] = np.stack([rng.uniform(0, 640, n_out), rng.uniform(0, 480, n_out)], 1).astype(np.float32)
    return X, uv


def main():
    rng = np.random.default_rng(4)
    cases = []
    # single-primitive known answers
    prim = []
    for _ in range(6):
        X, uv = make_case(rng, 6, 0.0, 0.5)
        ok, r, t = cv2.solvePnP(X, uv, K, None, flags=cv2.SOLVEPNP_EPNP)
        prim.append({"kind": "epnp6", "X": X.tolist(), "uv": uv.tolist(), "rvec": r.ravel().tolist(), "tvec": t.ravel().tolist()})
    for n in (12, 60, 300):
        X, uv = make_case(rng, n, 0.0, 0.7)
        ok, r0, t0 = cv2.solvePnP(X[:6], uv[:6], K, None, flags=cv2.SOLVEPNP_EPNP)
        ok, r, t = cv2.solvePnP(X, uv, K, None, r0.copy(), t0.copy(), True, cv2.SOLVEPNP_ITERATIVE)
        prim.append({"kind": "iterative", "X": X.tolist(), "uv": uv.tolist(), "rvec0": r0.ravel().tolist(), "tvec0": t0.ravel().tolist(),
                     "rvec": r.ravel().tolist(), "tvec": t.ravel().tolist()})
    for (n, of, noise, refine) in [(40, 0.2, 0.5, 1), (200, 0.3, 0.5, 1), (600, 0.3, 0.5, 1), (1000, 0.3, 0.5, 1), (1000, 0.5, 0.8, 1),
                                   (300, 0.3, 0.5, 0), (150, 0.7, 0.5, 1), (25, 0.1, 0.3, 1), (500, 0.3, 0.5, 5)]:
        X, uv = make_case(rng, n, of, noise)
        ok, r, t, inl, iters = util3d_solve_pnp_ransac(X, uv, 300, 2.0, 20, refine)
        cases.append({"n": n, "iterations": 300, "reproj": 2.0, "min_inliers": 20, "refine": refine, "X": X.tolist(), "uv": uv.tolist(),
                      "ok": bool(ok), "rvec": None if r is None else np.asarray(r).ravel().tolist(),
                      "tvec": None if t is None else np.asarray(t).ravel().tolist(), "inliers": inl, "iterations_run": iters})
        print(n, of, refine, ok, len(inl), iters)
    OUT.write_text(json.dumps({"K": [525.0, 525.0, 320.0, 240.0], "opencv": cv2.__version__, "primitives": prim, "ransac": cases}))
    print("wrote", OUT, OUT.stat().st_size)


if __name__ == "__main__":
    main()
