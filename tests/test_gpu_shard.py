"""Word-range sharding through the C ABI on ONE GPU: two engines each own half of the vocabulary rows and posting lists, the
collectives are emulated with torch ops on the same device (all-gather = concatenation, all-reduce = sum).  Both stage-2 variants
(replicated resolve, and resolve sharded by frame + scoring from all-gathered word ids) must reproduce the single-engine result:
identical word ids, identical likelihood (the scores are exact fixed-point sums, so the shard order cannot change them)."""
import numpy as np
import pytest

from rtabmap_b200 import Engine, sharding, synth

pytestmark = pytest.mark.gpu

W, S, F, B = 6000, 300, 256, 6


def _world():
    vocab = synth.make_binary_vocabulary(W, 32, 21)
    ids = (np.arange(1, W + 1, dtype=np.int32) * 2)
    m = synth.make_map(ids, S, F, seed=2)
    q, _ = synth.make_query_frames(vocab, ids, m, B, F, seed=5)
    q = np.ascontiguousarray(q).reshape(B, F, 32)
    q[1, 200:] = 0          # a short frame: 200 valid descriptors, zero padding rows
    return vocab, ids, m, q


@pytest.mark.parametrize("by_frame", [False, True])
def test_two_word_range_shards_equal_one_engine(by_frame):
    import torch

    vocab, ids, m, q = _world()
    n_valid = np.full(B, F, np.int32)
    n_valid[1] = 200
    last = int(ids.max())
    full = Engine()
    full.add_words(ids, vocab)
    full.last_word_id = last
    full.update()
    full.load_csr(m.word_ids, m.row_ptr, m.sig, m.cnt)
    full.set_ni(m.sig_ids, m.ni)
    want_w, want_l = [], []
    for b in range(B):   # per-frame calls: the reference for short frames is the frame with only its valid descriptors
        w_, l_ = full.localize_batch(q[b, :n_valid[b]], 1, m.sig_ids, S + 1)
        want_w.append(w_[0])
        want_l.append(l_[0])

    G = 2
    shards = []
    for r in range(G):
        r0, r1 = sharding.shard_rows(W, G, r)
        e = Engine()
        e.add_words(ids[r0:r1], vocab[r0:r1])
        e.last_word_id = last
        e.update()
        e.shard_set_row_offset(r0)
        e.load_csr(*sharding.shard_csr(m.word_ids, m.row_ptr, m.sig, m.cnt, ids[r0:r1]))
        e.set_ni(m.sig_ids, m.ni)
        shards.append(e)

    nq = B * F
    d_q = torch.from_numpy(q.reshape(-1)).cuda()
    d_rowids = torch.from_numpy(ids).cuda()
    d_sig = torch.from_numpy(m.sig_ids.astype(np.int32)).cuda()
    d_nv = torch.from_numpy(n_valid).cuda()
    torch.cuda.synchronize()
    keys = []
    for e in shards:
        k = torch.zeros(nq * 2, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        e.shard_knn2_keys_dev(d_q.data_ptr(), nq, k.data_ptr())
        e.synchronize()
        keys.append(k)
    keys_all = torch.cat(keys)          # "all-gather"
    scores = []
    if not by_frame:
        words = []
        for e in shards:
            w_ = torch.zeros(nq, dtype=torch.int32, device="cuda")
            sc = torch.zeros(B * S, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()
            e.shard_resolve_score_dev(d_q.data_ptr(), B, F, keys_all.data_ptr(), G, d_rowids.data_ptr(), W, last, d_sig.data_ptr(), S, S + 1,
                                      w_.data_ptr(), sc.data_ptr())
            e.synchronize()
            words.append(w_)
            scores.append(sc)
        assert torch.equal(words[0], words[1])      # the replicated pass is deterministic
        words_all = words[0]
    else:
        parts = []
        for r, e in enumerate(shards):
            f0, f1 = sharding.shard_rows(B, G, r)
            w_ = torch.zeros((f1 - f0) * F, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            e.shard_resolve_frames_dev(d_q.data_ptr(), f0, f1 - f0, B, F, keys_all.data_ptr(), G, d_rowids.data_ptr(), last, d_nv.data_ptr(),
                                       w_.data_ptr())
            e.synchronize()
            parts.append(w_)
        words_all = torch.cat(parts)    # "all-gather" of the word ids
        for e in shards:
            sc = torch.zeros(B * S, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()
            e.shard_score_ids_dev(words_all.data_ptr(), B, F, d_sig.data_ptr(), S, S + 1, sc.data_ptr())
            e.synchronize()
            scores.append(sc)
    total = scores[0] + scores[1]       # "all-reduce(sum)"
    like = torch.zeros(B * S, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    shards[0].shard_finalize_dev(total.data_ptr(), B * S, like.data_ptr())
    shards[0].synchronize()
    got_w = words_all.cpu().numpy().reshape(B, F)
    got_l = like.cpu().numpy().reshape(B, S)
    for b in range(B):
        if by_frame or n_valid[b] == F:   # the replicated variant has no per-frame counts: short frames are the by-frame variant's job
            assert np.array_equal(got_w[b, :n_valid[b]], want_w[b]), f"frame {b}"
            assert np.allclose(got_l[b], want_l[b], rtol=1e-6, atol=1e-7), f"frame {b}"
        if by_frame:
            assert not got_w[b, n_valid[b]:].any()


@pytest.mark.parametrize("score_parts", [1, 2])
def test_fused_sharded_step_with_a_single_rank_communicator_equals_the_unsharded_call(score_parts, monkeypatch):
    """score_parts: the TF-IDF stage whole (what 1-2 ranks run) or in two halves with a reduce-scatter each (what 4+ ranks run).
    lcd_shard_process_frames_dev (the exchanges inside the library, NCCL loaded at run time) on a communicator of ONE rank: every
    collective degenerates to a copy, the two-half pipeline and all the bookkeeping still run, and the result must equal
    lcd_process_frames_dev on the same engine — word ids, likelihood, hypotheses, verification."""
    import torch

    monkeypatch.setenv("LCD_SHARD_SCORE_PARTS", str(score_parts))
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
    imgs, deps, places = synth.make_view_frames(world, 5, seed=4)   # an odd count: halves of 3 and 2 frames
    vp = Engine.verify_params(synth.CAMERA_K4, image_size=(320, 240))
    nf, ns = len(imgs), len(sm.sig_ids)
    d_img = torch.from_numpy(imgs).cuda()
    d_dep = torch.from_numpy(deps.view(np.int16)).cuda()
    d_sig = torch.from_numpy(sm.sig_ids).cuda()
    d_rowids = torch.from_numpy(world.word_ids).cuda()
    ext = torch.cuda.ExternalStream(eng.stream)
    with torch.cuda.stream(ext):
        w0 = torch.zeros(nf * 300, dtype=torch.int32, device="cuda")
        l0 = torch.zeros(nf * ns, dtype=torch.float32, device="cuda")
        eng.process_frames_dev(d_img.data_ptr(), nf, 320, 240, 3, d_dep.data_ptr(), 1, op, d_sig.data_ptr(), ns, ns + 1, vp, w0.data_ptr(), l0.data_ptr(),
                               True, 0.8, True)
        hyp0, res0 = eng.process_fetch(nf)
        eng.shard_comm_init(Engine.shard_unique_id(), 0, 1)
        w1 = torch.zeros(nf * 300, dtype=torch.int32, device="cuda")
        l1 = torch.zeros(nf * ns, dtype=torch.float32, device="cuda")
        for _ in range(2):  # twice: the exchange buffers and events are reused
            eng.shard_process_frames_dev(d_img.data_ptr(), nf, 320, 240, 3, d_dep.data_ptr(), 1, op, d_sig.data_ptr(), ns, ns + 1, d_rowids.data_ptr(),
                                         int(world.word_ids.max()), vp, w1.data_ptr(), l1.data_ptr(), True, 0.8, True)
        hyp1, res1 = eng.process_fetch(nf)
        eng.shard_comm_destroy()
    assert torch.equal(w0, w1)
    assert torch.equal(l0, l1)
    assert np.array_equal(hyp0, hyp1) and (hyp0 > 0).all()
    for a, b in zip(res0, res1):
        assert a["ok"] == b["ok"] and a["n_inliers"] == b["n_inliers"] and np.array_equal(a["rvec"], b["rvec"])
