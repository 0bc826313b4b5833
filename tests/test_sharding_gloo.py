"""World-size-2 test of the word-range sharding logic over the gloo backend (CPU only).

Each rank owns half of the vocabulary rows and half of the posting lists.  The per-rank arithmetic is done with
the CPU oracle here (the GPU kernels are checked in test_gpu_parity.py); what this test pins is the HOST logic of
the N>1 path (rtabmap_b200/sharding.py + the collective pattern bench.py uses): packed (dist<<22 | global row) keys
all-gathered and merged must equal the unsharded top-2, and the all-reduced per-shard scores must equal the
unsharded likelihood."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_py as orc
from rtabmap_b200 import sharding, synth

W, S, F, Q = 2000, 120, 80, 96


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vocab = synth.make_binary_vocabulary(W, 32, 1)
    ids = np.arange(1, W + 1, dtype=np.int32) * 3          # non-contiguous ids
    m = synth.make_map(ids, S, F, seed=2)
    q, _ = synth.make_query_frames(vocab, ids, m, 1, Q, seed=3)

    r0, r1 = sharding.shard_rows(W, world, rank)
    d = orc.OracleDictionary(incremental=True)
    d.add_words(ids[r0:r1], vocab[r0:r1])
    d.update()
    w_, p_, s_, c_ = sharding.shard_csr(m.word_ids, m.row_ptr, m.sig, m.cnt, ids[r0:r1])
    d.load_csr(w_, p_, s_, c_)
    d.set_ni(m.sig_ids, m.ni)

    # stage 1: local top-2 -> packed keys with GLOBAL rows -> all-gather -> merge
    idx, dd = orc.knn2_raw(vocab[r0:r1], q)
    keys = np.where(idx >= 0, sharding.pack_keys(np.maximum(dd, 0), np.maximum(idx, 0) + r0), sharding.KEY_NONE).astype(np.uint32)
    gathered = [torch.zeros(Q * 2, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(keys.reshape(-1).view(np.int32).copy()))
    allk = np.stack([g.numpy().view(np.uint32).reshape(Q, 2) for g in gathered])
    merged = sharding.merge_top2(allk)

    # stage 2: each rank scores the words it owns; all-reduce(sum) of exact fixed-point sums
    best_row = (merged[:, 0] & sharding.KEY_ROW_MASK).astype(np.int64)
    words = ids[best_row]                                   # fixed-dictionary style assignment is enough for this test
    lik = d.likelihood(words, m.sig_ids, S + 1)             # words this rank does not own contribute nothing
    fx = torch.from_numpy(np.rint(lik.astype(np.float64) * 2.0 ** 40).astype(np.int64))
    dist.all_reduce(fx, op=dist.ReduceOp.SUM)
    if rank == 0:
        out["merged"] = merged
        out["lik"] = fx.numpy().astype(np.float64) / 2.0 ** 40
        out["words"] = words
    dist.destroy_process_group()


def test_word_range_sharding_over_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        merged, lik, words = out["merged"], out["lik"], out["words"]
    # unsharded truth
    vocab = synth.make_binary_vocabulary(W, 32, 1)
    ids = np.arange(1, W + 1, dtype=np.int32) * 3
    m = synth.make_map(ids, S, F, seed=2)
    q, _ = synth.make_query_frames(vocab, ids, m, 1, Q, seed=3)
    idx, dd = orc.knn2_raw(vocab, q)
    want = sharding.pack_keys(dd, idx)
    assert np.array_equal(merged, want)                     # same (distance, lowest row) order as the single index
    d = orc.OracleDictionary()
    d.add_words(ids, vocab)
    d.update()
    d.load_csr(m.word_ids, m.row_ptr, m.sig, m.cnt)
    full = d.likelihood(words, m.sig_ids, S + 1)
    assert np.allclose(lik, full, atol=1e-4, rtol=1e-4)


def test_shard_rows_partition_is_exact():
    for n, g in ((49152, 8), (10, 3), (7, 8), (1000, 1)):
        spans = [sharding.shard_rows(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


# ---- the frame-sharded stage 2 that bench.py runs at N > 1 --------------------------------------------------
BF, FF = 6, 64   # frames of the whole job, descriptors per frame


def _worker_frames(rank, world, port, out):
    """Every rank resolves ITS frames, the word ids are all-gathered in rank order, every rank scores ALL frames on its word
    range, and the exact fixed-point sums are reduced; each rank keeps the rows of its own frames."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vocab = synth.make_binary_vocabulary(W, 32, 1)
    ids = np.arange(1, W + 1, dtype=np.int32) * 3
    m = synth.make_map(ids, S, F, seed=2)
    q, _ = synth.make_query_frames(vocab, ids, m, BF, FF, seed=7)
    bl = BF // world                                   # frames per rank, rank r owns frames [r*bl, (r+1)*bl)
    f0 = rank * bl
    # resolve the local frames (the merged top-2 of the shards equals the full index, pinned by the test above)
    full = orc.OracleDictionary(incremental=True)
    full.add_words(ids, vocab)
    full.last_word_id = int(ids.max())
    full.update()
    mine = np.zeros((bl, FF), np.int32)
    for b in range(bl):
        mine[b], _ = full.localize_ro(q[(f0 + b) * FF:(f0 + b + 1) * FF], m.sig_ids, S + 1, want_like=False)
    gathered = [torch.zeros(bl * FF, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine.reshape(-1).copy()))
    words_all = torch.cat(gathered).numpy().reshape(BF, FF)
    # score every frame on this rank's word range
    r0, r1 = sharding.shard_rows(W, world, rank)
    d = orc.OracleDictionary(incremental=True)
    d.add_words(ids[r0:r1], vocab[r0:r1])
    d.update()
    d.load_csr(*sharding.shard_csr(m.word_ids, m.row_ptr, m.sig, m.cnt, ids[r0:r1]))
    d.set_ni(m.sig_ids, m.ni)
    part = np.stack([d.likelihood(words_all[b], m.sig_ids, S + 1) for b in range(BF)])
    fx = torch.from_numpy(np.rint(part.astype(np.float64) * 2.0 ** 40).astype(np.int64))
    dist.all_reduce(fx, op=dist.ReduceOp.SUM)          # gloo has no reduce-scatter: reduce, then keep the local rows
    out[rank] = (words_all, fx.numpy()[f0:f0 + bl].astype(np.float64) / 2.0 ** 40)
    dist.destroy_process_group()


def test_frame_sharded_stage2_over_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_frames, args=(world, _free_port(), out), nprocs=world, join=True)
        res = {r: out[r] for r in range(world)}
    vocab = synth.make_binary_vocabulary(W, 32, 1)
    ids = np.arange(1, W + 1, dtype=np.int32) * 3
    m = synth.make_map(ids, S, F, seed=2)
    q, _ = synth.make_query_frames(vocab, ids, m, BF, FF, seed=7)
    d = orc.OracleDictionary(incremental=True)
    d.add_words(ids, vocab)
    d.last_word_id = int(ids.max())
    d.update()
    d.load_csr(m.word_ids, m.row_ptr, m.sig, m.cnt)
    d.set_ni(m.sig_ids, m.ni)
    bl = BF // world
    assert np.array_equal(res[0][0], res[1][0])                      # every rank sees the same all-gathered word ids
    for b in range(BF):
        w, _ = d.localize_ro(q[b * FF:(b + 1) * FF], m.sig_ids, S + 1, want_like=False)
        assert np.array_equal(res[0][0][b], w)
        # the sharded sums against the unsharded TF-IDF of the same word ids (the per-frame self reference of the
        # localisation mode is an engine detail covered by tests/test_gpu_shard.py; this test pins the partitioning)
        assert np.allclose(res[b // bl][1][b % bl], d.likelihood(w, m.sig_ids, S + 1), atol=1e-4, rtol=1e-4)
