"""Parity of the CUDA ORB (detect + describe + 3-D lifting) against OpenCV's cv::ORB — the third-party
implementation RTAB-Map's Feature2D calls (Features2d.cpp:1614/1657/1714) — through the oracle's restated
RTAB-Map wrapper logic (oracle/feature2d_py.py).  Bar: identical keypoints IN THE SAME ORDER (pt, octave,
size bit-exact; response and angle bit-exact floats), identical descriptor bytes, identical 3-D points."""
import numpy as np
import pytest

from oracle import feature2d_py as f2d
from rtabmap_b200 import Engine, synth

pytestmark = pytest.mark.gpu
K4 = synth.CAMERA_K4


def check_frame(got, want, exact_desc=True):
    kp_g, d_g, x_g = got
    kp_w, d_w, x_w = want
    assert len(kp_g) == len(kp_w), (len(kp_g), len(kp_w))
    assert np.array_equal(kp_g[:, [0, 1, 2, 5]], kp_w[:, [0, 1, 2, 5]]), "keypoint set / order differs"
    assert np.array_equal(kp_g[:, 4].view(np.uint32), kp_w[:, 4].view(np.uint32)), "Harris responses differ"
    assert np.array_equal(kp_g[:, 3].view(np.uint32), kp_w[:, 3].view(np.uint32)), "angles differ"
    bad_bits = int(np.unpackbits(d_g ^ d_w).sum())
    if exact_desc:
        assert bad_bits == 0, f"{bad_bits} descriptor bits differ"
    assert np.array_equal(np.isnan(x_g), np.isnan(x_w))
    assert np.array_equal(np.nan_to_num(x_g).view(np.uint32), np.nan_to_num(x_w).view(np.uint32)), "3-D points differ"
    return bad_bits


@pytest.mark.parametrize("seed", [5, 6, 7])
def test_orb_gray_with_depth_mask(seed):
    img = synth.make_image(480, 640, seed)
    depth = synth.make_depth(480, 640, seed + 10)
    eng = Engine()
    p = f2d.OrbParams()
    got = eng.orb_detect_describe(img[None], depth[None], Engine.orb_params(K4))[0]
    want = f2d.detect_describe(img, depth, K4, p)
    assert len(want[0]) > 900
    check_frame(got, want)


def test_orb_bgr_float_depth_batch():
    imgs = np.stack([synth.make_image(480, 640, 20 + i, bgr=True) for i in range(3)])
    depth = np.stack([synth.make_depth(480, 640, 30 + i, as_float=True) for i in range(3)])
    eng = Engine()
    got = eng.orb_detect_describe(imgs, depth, Engine.orb_params(K4, min_depth=0.5, max_depth=3.5))
    p = f2d.OrbParams(min_depth=0.5, max_depth=3.5)
    for i in range(3):
        check_frame(got[i], f2d.detect_describe(imgs[i], depth[i], K4, p))


def test_orb_no_depth_and_few_features():
    img = synth.make_image(240, 320, 3, n_rects=40)          # low texture: fewer corners than the quota
    eng = Engine()
    got = eng.orb_detect_describe(img[None], None, Engine.orb_params(K4, n_features=400, depth_as_mask=False))[0]
    want = f2d.detect_describe(img, None, K4, f2d.OrbParams(n_features=400, depth_as_mask=False))
    check_frame(got, want)
    assert np.isnan(got[2]).all()


def test_orb_limit_keypoints_when_levels_overflow_quota():
    # many equal Harris responses are unlikely; force the limit path with a smaller Kp/MaxFeatures than cv::ORB's own
    img = synth.make_image(480, 640, 11)
    eng = Engine()
    for nf, lv in ((500, 3), (300, 2), (1000, 1), (800, 4)):
        got = eng.orb_detect_describe(img[None], None, Engine.orb_params(K4, n_features=nf, n_levels=lv, depth_as_mask=False))[0]
        want = f2d.detect_describe(img, None, K4, f2d.OrbParams(n_features=nf, n_levels=lv, depth_as_mask=False))
        check_frame(got, want)


def test_orb_rejects_unsupported_parameters():
    from rtabmap_b200 import LcdError

    eng = Engine()
    img = np.zeros((1, 480, 640), np.uint8)
    with pytest.raises(LcdError):
        eng.orb_detect_describe(img, None, Engine.orb_params(K4, scale_factor=1.2))
    with pytest.raises(LcdError):
        eng.orb_detect_describe(np.zeros((1, 481, 641), np.uint8), None, Engine.orb_params(K4))
