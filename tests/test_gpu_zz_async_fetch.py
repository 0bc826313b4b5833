"""lcd_process_fetch_async: the download of a step's hypotheses and verification results queued behind the step (what a caller with
several steps in flight uses) must deliver what the synchronising lcd_process_fetch delivers."""
import ctypes

import numpy as np
import pytest

from rtabmap_b200 import Engine, synth
from rtabmap_b200.capi import VerifyResult

pytestmark = pytest.mark.gpu


def test_async_download_equals_the_synchronous_one():
    import torch

    eng = Engine(max_words=8192, max_signatures=400)
    op = Engine.orb_params(synth.CAMERA_K4, n_features=300)
    world = synth.make_place_world(lambda im, dp: eng.orb_detect_describe(im[None], dp[None], op, cap=300)[0], 6, 1500, 120, 300, 240, 320)
    sm = world.smap
    eng.add_words(world.word_ids, world.vocab)
    eng.last_word_id = int(world.word_ids.max())
    eng.update()
    eng.load_csr(sm.word_ids, sm.row_ptr, sm.sig, sm.cnt)
    eng.set_ni(sm.sig_ids, sm.ni)
    eng.sig_add_batch(sm.sig_ids, world.store.desc, world.store.xyz, sm.ni)
    imgs, deps, places = synth.make_view_frames(world, 4, seed=4)
    vp = Engine.verify_params(synth.CAMERA_K4, image_size=(320, 240))
    nf, ns = len(imgs), len(sm.sig_ids)
    d_img = torch.from_numpy(imgs).cuda()
    d_dep = torch.from_numpy(deps.view(np.int16)).cuda()
    d_sig = torch.from_numpy(sm.sig_ids).cuda()
    ext = torch.cuda.ExternalStream(eng.stream)
    with torch.cuda.stream(ext):
        w0 = torch.zeros(nf * 300, dtype=torch.int32, device="cuda")
        l0 = torch.zeros(nf * ns, dtype=torch.float32, device="cuda")
        eng.process_frames_dev(d_img.data_ptr(), nf, 320, 240, 3, d_dep.data_ptr(), 1, op, d_sig.data_ptr(), ns, ns + 1, vp, w0.data_ptr(), l0.data_ptr(),
                               True, 0.8, True)
        hyp_pin = torch.zeros(nf, dtype=torch.int32).pin_memory()
        res_pin = torch.zeros(ctypes.sizeof(VerifyResult) * nf, dtype=torch.uint8).pin_memory()
        eng.process_fetch_async(nf, hyp_pin.data_ptr(), res_pin.data_ptr(), eng.stream)
        ext.synchronize()
        got = Engine.results_from_buffer(res_pin.numpy(), nf)
        hyp, want = eng.process_fetch(nf)
    assert np.array_equal(hyp_pin.numpy(), hyp)
    for a, b in zip(want, got):
        assert a["ok"] == b["ok"] and a["n_matches"] == b["n_matches"] and a["n_inliers"] == b["n_inliers"] and a["iterations_run"] == b["iterations_run"]
        assert np.array_equal(a["rvec"], b["rvec"]) and np.array_equal(a["tvec"], b["tvec"]) and np.array_equal(a["transform"], b["transform"])
        assert np.array_equal(a["covariance"], b["covariance"])
