"""Parity of the verification kernels (pair matching + PnP/RANSAC + refinement) against the CPU oracle
and against the OpenCV-derived golden answers (tests/golden/pnp_golden.json).
Bit-exact: word ids of the matching, correspondences, inlier sets, iteration counts.
Within 1e-4 (north_star tolerance on pose): rvec / tvec / transform (observed ~1e-9)."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle_py as orc
from rtabmap_b200 import Engine, synth

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).parent / "golden" / "pnp_golden.json").read_text())
K4 = GOLD["K"]
POSE_TOL = 1e-4


def make_pair(rng, n, flip=0.04, outlier_frac=0.3, nan_every=50, dup_every=0):
    import cv2  # only to synthesise the geometry (projectPoints); the checker is the oracle

    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    if dup_every:
        desc[dup_every::dup_every] = synth.flip_bits(desc[0:1], 0.01, rng)  # near-duplicates inside FROM -> non-unique words
    X = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(0.5, 5, n)], 1).astype(np.float32)
    if nan_every:
        X[::nan_every] = np.nan
    rv = rng.normal(0, 0.08, 3)
    tv = rng.normal(0, 0.15, 3)
    uv, _ = cv2.projectPoints(np.nan_to_num(X).astype(np.float64), rv, tv, K, None)
    uv = (uv.reshape(-1, 2) + rng.normal(0, 0.5, (n, 2))).astype(np.float32)
    perm = rng.permutation(n)
    desc_to = synth.flip_bits(desc[perm], flip, rng)
    uv_to = uv[perm].copy()
    bad = rng.permutation(n)[: int(outlier_frac * n)]
    uv_to[bad] = np.stack([rng.uniform(0, 640, len(bad)), rng.uniform(0, 480, len(bad))], 1)
    return desc, X, desc_to, uv_to


def test_match_pairs_bit_exact():
    rng = np.random.default_rng(11)
    eng = Engine()
    pairs = [make_pair(rng, n, dup_every=d) for n, d in [(300, 0), (300, 7), (257, 0), (64, 5)]]
    cap = 300
    F = np.zeros((len(pairs), cap, 32), np.uint8)
    T = np.zeros((len(pairs), cap, 32), np.uint8)
    nf = []
    for i, (df, X, dt, uv) in enumerate(pairs):
        F[i, :len(df)] = df
        T[i, :len(dt)] = dt
        nf.append(len(df))
    fid, tid = eng.match_pairs(F, T, nf, nf)
    for i, (df, X, dt, uv) in enumerate(pairs):
        f_o, t_o = orc.match_pair(df, dt)
        assert np.array_equal(fid[i, :nf[i]], f_o) and np.array_equal(tid[i, :nf[i]], t_o)


@pytest.mark.parametrize("case", GOLD["ransac"], ids=lambda c: f"n{c['n']}_refine{c['refine']}")
def test_pnp_ransac_matches_opencv_golden(case):
    """Known answers computed with real OpenCV primitives: feed the correspondences through unique descriptors."""
    n = case["n"]
    rng = np.random.default_rng(n)
    desc = rng.integers(0, 256, (1, n, 32), dtype=np.uint8)
    eng = Engine()
    r = eng.verify_batch(desc, np.asarray(case["X"], np.float32)[None], desc, np.asarray(case["uv"], np.float32)[None], K4,
                         min_inliers=case["min_inliers"], iterations=case["iterations"], reproj_error=case["reproj"],
                         refine_iterations=case["refine"])[0]
    assert r["matches"].tolist() == list(range(1, n + 1))
    assert r["iterations_run"] == case["iterations_run"]
    assert (r["inliers"] - 1).tolist() == case["inliers"]
    assert np.allclose(r["rvec"], case["rvec"], atol=1e-6) and np.allclose(r["tvec"], case["tvec"], atol=1e-6)


def test_verify_batch_matches_oracle():
    rng = np.random.default_rng(21)
    specs = [(1000, 0.3), (1000, 0.5), (400, 0.2), (120, 0.6), (30, 0.1), (15, 0.0)]
    cap = 1000
    B = len(specs)
    F = np.zeros((B, cap, 32), np.uint8); T = np.zeros((B, cap, 32), np.uint8)
    X = np.full((B, cap, 3), np.nan, np.float32); U = np.zeros((B, cap, 2), np.float32)
    ns = []
    raw = []
    for i, (n, of) in enumerate(specs):
        df, x, dt, uv = make_pair(rng, n, outlier_frac=of)
        F[i, :n], T[i, :n], X[i, :n], U[i, :n] = df, dt, x, uv
        ns.append(n)
        raw.append((df, x, dt, uv))
    eng = Engine()
    out = eng.verify_batch(F, X, T, U, K4, ns, ns)
    for i, (df, x, dt, uv) in enumerate(raw):
        o = orc.verify_pair(df, x, dt, uv, K4)
        g = out[i]
        assert g["ok"] == o["ok"], i
        assert np.array_equal(g["matches"], o["matches"]), i
        assert np.array_equal(g["inliers"], o["inliers"]), i
        assert np.allclose(g["rvec"], o["rvec"], atol=POSE_TOL) and np.allclose(g["tvec"], o["tvec"], atol=POSE_TOL)
        assert np.abs(g["rvec"] - o["rvec"]).max() < 1e-7 and np.abs(g["tvec"] - o["tvec"]).max() < 1e-7
        assert np.allclose(g["transform"], o["transform"], atol=POSE_TOL)
    assert out[0]["ok"] and not out[5]["ok"]  # 15 correspondences < Vis/MinInliers


def test_verify_rejects_unrelated_pair():
    rng = np.random.default_rng(5)
    df, x, dt, uv = make_pair(rng, 500)
    other = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    eng = Engine()
    r = eng.verify_batch(df[None], x[None], other[None], uv[None], K4)[0]
    o = orc.verify_pair(df, x, other, uv, K4)
    assert not r["ok"] and not o["ok"] and len(r["matches"]) == len(o["matches"])


def test_process_batch_matches_oracle():
    """Fused quantise -> score -> arg-max hypothesis -> verify against the device-resident signature store."""
    vocab = synth.make_binary_vocabulary(6000, 32, 1)
    ids = np.arange(1, 6001, dtype=np.int32)
    m = synth.make_map(ids, 250, 400, seed=2)
    st = synth.make_signature_store(vocab, ids, m)
    B, F = 5, 400
    q, uv, places, poses = synth.make_query_frames_geo(st, m, B)
    eng = Engine()
    o = orc.OracleDictionary()
    for d in (eng, o):
        d.add_words(ids, vocab)
        d.last_word_id = 6000
        d.update()
        d.load_csr(m.word_ids, m.row_ptr, m.sig, m.cnt)
    eng.sig_add_batch(m.sig_ids[:100], st.desc[:100], st.xyz[:100])
    eng.sig_add_batch(m.sig_ids[100:], st.desc[100:], st.xyz[100:])
    assert eng.sig_count() == 250
    vp = Engine.verify_params(synth.CAMERA_K4)
    words, like, hyp, res = eng.process_batch(q, uv, B, m.sig_ids, 251, vp)
    for b in range(B):
        fq, fuv = q[b * F:(b + 1) * F], uv[b * F:(b + 1) * F]
        w_o, l_o = o.localize_ro(fq, m.sig_ids, 251)
        assert np.array_equal(words[b], w_o)
        assert np.allclose(like[b], l_o, atol=1e-4, rtol=1e-4)
        h = int(m.sig_ids[np.argmax(l_o)])
        assert hyp[b] == h == places[b]
        v = orc.verify_pair(st.desc[h - 1], st.xyz[h - 1], fq, fuv, synth.CAMERA_K4)
        r = res[b]
        assert r["ok"] == v["ok"] and r["n_matches"] == len(v["matches"]) and r["n_inliers"] == len(v["inliers"])
        assert np.allclose(r["rvec"], v["rvec"], atol=POSE_TOL) and np.allclose(r["tvec"], v["tvec"], atol=POSE_TOL)
        assert np.allclose(r["transform"], v["transform"], atol=POSE_TOL)
        assert r["ok"] and np.abs(r["rvec"] - poses[b, :3]).max() < 5e-3 and np.abs(r["tvec"] - poses[b, 3:]).max() < 1e-2
    # a removed signature can no longer be verified
    eng.sig_remove(int(places[0]))
    _, _, hyp2, res2 = eng.process_batch(q[:F], uv[:F], 1, m.sig_ids, 251, vp)
    assert hyp2[0] == places[0] and not res2[0]["ok"] and res2[0]["n_matches"] == 0
