"""Mapping mode (Memory::update + computeLikelihood frame after frame): the pipelined lcd_map_detect_async / lcd_map_frame pair against
(a) the same work issued call by call (lcd_orb_detect_describe, lcd_dict_update, lcd_dict_quantize, lcd_index_score) and (b) cv::ORB +
the CPU oracle.  Bit-exact: keypoints, descriptors, word ids, dictionary growth; likelihood 1e-4 (exact sums on the GPU side)."""
import numpy as np
import pytest

from oracle import feature2d_py as f2d
from oracle import oracle_py as orc
from rtabmap_b200 import Engine, synth

pytestmark = pytest.mark.gpu


def test_mapping_stream_equals_per_call_sequence_and_oracle():
    rng = np.random.default_rng(3)
    n_places, n_frames, stm = 5, 22, 4
    K4 = synth.CAMERA_K4
    op = Engine.orb_params(K4, n_features=500)
    base_img = [synth.make_image(240, 320, 900 + p, bgr=True) for p in range(n_places)]
    base_dep = [synth.make_depth(240, 320, 950 + p) for p in range(n_places)]
    frames = []
    for t in range(n_frames):  # the robot goes round the places: revisits from the second lap on
        p = t % n_places
        frames.append(synth.render_view(base_img[p], base_dep[p], int(rng.integers(-4, 5)), int(rng.integers(-4, 5)), 2.0, rng))
    pipe, seq = Engine(), Engine()
    o = orc.OracleDictionary(0, 32, True, 0.8, True)
    pipe.map_detect_async(frames[0][0], frames[0][1], op)
    revisit_hits = 0
    for t in range(n_frames):
        sig = t + 1
        if t + 1 < n_frames:
            pipe.map_detect_async(frames[t + 1][0], frames[t + 1][1], op)   # detection of the next frame overlaps this frame's update
        wm = np.arange(1, sig - stm + 1, dtype=np.int32)
        n_kp, words, n_new, like, kp, desc, xyz = pipe.map_frame(sig, wm, sig, want_features=True)
        # (a) call by call on a second engine
        kp_s, desc_s, xyz_s = seq.orb_detect_describe(frames[t][0][None], frames[t][1][None], op, cap=500)[0]
        seq.update()
        w_s, n_new_s = seq.quantize(desc_s, sig)
        assert n_kp == len(kp_s) and np.array_equal(desc, desc_s) and np.array_equal(xyz, xyz_s, equal_nan=True)
        assert np.array_equal(words, w_s) and n_new == n_new_s, f"frame {t}"
        # (b) cv::ORB + oracle
        kp_o, d_o, x_o = f2d.detect_describe(frames[t][0], frames[t][1], K4, f2d.OrbParams(n_features=500))
        assert np.array_equal(desc, d_o) and np.array_equal(xyz, x_o, equal_nan=True)
        o.update()
        w_o = o.add_new_words(d_o, sig)
        assert np.array_equal(words, w_o), f"frame {t}: word ids differ from the oracle"
        if len(wm):
            l_s = seq.score(w_s, wm, sig)
            l_o = o.likelihood(w_o, wm, sig)
            assert np.array_equal(like, l_s)
            assert np.allclose(like, l_o, rtol=1e-4, atol=1e-6)
            if t >= n_places + stm:
                revisit_hits += int((int(np.argmax(like)) % n_places) == (t % n_places))
    assert pipe.size() == seq.size() == o.size() and pipe.last_word_id == o.last_word_id
    assert revisit_hits >= (n_frames - n_places - stm) - 1   # the TF-IDF arg-max is the revisited place


def test_map_frame_without_a_detection_fails_loudly():
    eng = Engine()
    from rtabmap_b200.capi import LcdError

    with pytest.raises(LcdError):
        eng.map_frame(1)
