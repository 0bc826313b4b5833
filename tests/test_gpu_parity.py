"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bit-exact: word ids, new-word ids, nearest-neighbour ids/distances, inverted-index contents.
Within 1e-4 (north_star tolerance for float work): likelihoods.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle_py as orc
from rtabmap_b200 import Engine, VWDictionaryB200, synth

pytestmark = pytest.mark.gpu

LIK_TOL = 1e-4  # north_star: "within 1e-4 on likelihood"


def assert_likelihood_close(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    err = np.abs(got - want)
    assert np.all(err <= LIK_TOL * np.maximum(1.0, np.abs(want))), f"max err {err.max()}"


def make_pair(n_words, dim=32, seed=1, incremental=True, nndr=0.8, cmp_new=True, id_stride=1):
    vocab = synth.make_binary_vocabulary(n_words, dim, seed)
    ids = (np.arange(n_words, dtype=np.int32) * id_stride + 1)
    eng = Engine(desc_dim=dim)
    o = orc.OracleDictionary(0, dim, incremental, nndr, cmp_new)
    if n_words:
        eng.add_words(ids, vocab)
        o.add_words(ids, vocab)
    last = int(ids.max()) if n_words else 0
    eng.last_word_id = last
    o.last_word_id = last
    eng.update()
    o.update()
    return eng, o, vocab, ids


# ---------------------------------------------------------------- knn2 -------------------------
@pytest.mark.parametrize("n_words", [0, 1, 2, 3, 33, 1000, 5000])
def test_knn2_bit_exact(n_words):
    eng, o, vocab, ids = make_pair(n_words, id_stride=3)
    rng = np.random.default_rng(n_words + 11)
    q = rng.integers(0, 256, (257, 32), dtype=np.uint8)
    if n_words >= 33:
        q[:20] = synth.flip_bits(vocab[rng.integers(0, n_words, 20)], 0.05, rng)
        q[20] = vocab[7]
    g = eng.knn2(q)
    w = o.knn2(q)
    for a, b in zip(g, w):
        assert np.array_equal(a, b)


def test_knn2_ties_resolve_to_lowest_row():
    vocab = synth.make_binary_vocabulary(600, 32, 5)
    vocab[100] = vocab[3]
    vocab[400] = vocab[3]
    vocab[599] = vocab[3]
    ids = np.arange(1, 601, dtype=np.int32)
    eng = Engine()
    eng.add_words(ids, vocab)
    eng.update()
    id1, d1, id2, d2 = eng.knn2(vocab[3:4])
    assert (id1[0], d1[0], id2[0], d2[0]) == (4, 0.0, 101, 0.0)


@pytest.mark.parametrize("n_words,nq", [(1, 1), (255, 129), (256, 128), (257, 300), (4097, 1000), (20000, 2500)])
def test_nn_kernels_agree(n_words, nq):
    """The tensor-core kernel (s8 GEMM of the +-1 encoded bits) and the XOR/POPC kernel return the same two neighbours and
    distances for every query, including ties (lowest row wins), partly filled word / query tiles and duplicate words."""
    eng, o, vocab, ids = make_pair(n_words, id_stride=2, seed=n_words)
    rng = np.random.default_rng(n_words + nq)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    k = min(nq, n_words)
    q[:k] = synth.flip_bits(vocab[rng.integers(0, n_words, k)], 0.04, rng)
    q[0] = vocab[0]
    eng.nn_select(1)
    t = eng.knn2(q)
    assert eng.nn_last_kernel == 1
    eng.nn_select(0)
    p = eng.knn2(q)
    assert eng.nn_last_kernel == 0
    w = o.knn2(q)
    for a, b, c in zip(t, p, w):
        assert np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.parametrize("dim", [16, 64])
def test_knn2_other_descriptor_sizes(dim):
    eng, o, vocab, ids = make_pair(700, dim=dim, seed=9)
    rng = np.random.default_rng(3)
    q = synth.flip_bits(vocab[rng.integers(0, 700, 90)], 0.08, rng)
    for a, b in zip(eng.knn2(q), o.knn2(q)):
        assert np.array_equal(a, b)


# ---------------------------------------------------------------- addNewWords ------------------
@pytest.mark.parametrize("cmp_new", [True, False])
@pytest.mark.parametrize("nndr", [0.8, 0.6])
def test_incremental_stream_bit_exact(cmp_new, nndr):
    """Mapping mode from an EMPTY dictionary: update() + addNewWords() per frame, like Memory::update."""
    rng = np.random.default_rng(42)
    places = synth.make_binary_vocabulary(3000, 32, 77)
    eng = Engine()
    o = orc.OracleDictionary(0, 32, True, nndr, cmp_new)
    for t in range(1, 13):
        sel = rng.integers(0, 3000, 300)
        frame = synth.flip_bits(places[sel], 0.04, rng)
        frame[::17] = rng.integers(0, 256, (len(frame[::17]), 32), dtype=np.uint8)
        if t % 3 == 0:  # near-duplicate descriptors inside one frame -> intra-frame dependency chains
            frame[1::7] = synth.flip_bits(frame[0::7][: len(frame[1::7])], 0.02, rng)
        eng.update()
        o.update()
        g, n_new = eng.quantize(frame, t, True, nndr, cmp_new)
        w = o.add_new_words(frame, t)
        assert np.array_equal(g, w), f"frame {t}"
        assert eng.last_word_id == o.last_word_id
        assert eng.not_indexed_size() == o.not_indexed_size() == n_new
    eng.update()
    o.update()
    assert np.array_equal(eng.get_indexed()[0], o.indexed_ids())
    # inverted index contents, word by word
    for wid in rng.integers(1, o.last_word_id + 1, 200):
        s1, c1 = eng.get_refs(int(wid))
        s2, c2 = o.get_refs(int(wid))
        assert np.array_equal(s1, s2) and np.array_equal(c1, c2)


def test_fixed_dictionary_quantisation():
    eng, o, vocab, ids = make_pair(4000, incremental=False, id_stride=2)
    rng = np.random.default_rng(8)
    q = synth.flip_bits(vocab[rng.integers(0, 4000, 500)], 0.1, rng)
    g, n_new = eng.quantize(q, 5, False, 0.8, True)
    w = o.add_new_words(q, 5)
    assert n_new == 0 and np.array_equal(g, w)


def test_remove_words_and_signatures_then_continue():
    rng = np.random.default_rng(5)
    eng, o, vocab, ids = make_pair(1500)
    frames = [synth.flip_bits(vocab[rng.integers(0, 1500, 200)], 0.05, rng) for _ in range(6)]
    for t, f in enumerate(frames[:3], start=1):
        eng.update(); o.update()
        assert np.array_equal(eng.quantize(f, t)[0], o.add_new_words(f, t))
    # forget signature 2, then delete some now-unused + some arbitrary words (Memory::cleanUnusedWords)
    eng.remove_sig(2); o.remove_sig(2)
    victims = np.unique(np.concatenate([rng.integers(1, 1501, 40), np.arange(o.last_word_id - 3, o.last_word_id + 1)])).astype(np.int32)
    eng.remove_words(victims); o.remove_words(victims)
    for t, f in enumerate(frames[3:], start=4):
        eng.update(); o.update()
        assert np.array_equal(eng.get_indexed()[0], o.indexed_ids())
        assert np.array_equal(eng.quantize(f, t)[0], o.add_new_words(f, t))
    assert eng.size() == o.size()


def test_add_words_out_of_order_and_find_nn():
    rng = np.random.default_rng(6)
    eng, o, vocab, ids = make_pair(900, id_stride=5)
    extra = synth.make_binary_vocabulary(40, 32, 99)
    extra_ids = np.array([4502 + 5 * k for k in range(40)], np.int32)[::-1].copy()  # descending: set<int> re-sorts them
    eng.add_words(extra_ids, extra); o.add_words(extra_ids, extra)
    q = np.concatenate([synth.flip_bits(extra[:20], 0.03, rng), synth.flip_bits(vocab[:60], 0.05, rng),
                        rng.integers(0, 256, (30, 32), dtype=np.uint8)])
    assert np.array_equal(eng.find_nn(q, True, 0.8), o.find_nn(q))       # not-indexed words searched too
    eng.update(); o.update()
    assert np.array_equal(eng.get_indexed()[0], o.indexed_ids())
    assert np.array_equal(eng.find_nn(q, True, 0.8), o.find_nn(q))


# ---------------------------------------------------------------- scoring ----------------------
def test_tfidf_golden_vector_on_gpu():
    from golden_util import GOLD, load_golden_into

    remap = lambda s: 1000 if s == -1 else s
    eng = Engine()
    load_golden_into(eng, remap)
    ids = [remap(s) for s in GOLD["sig_ids"]]
    lik = eng.score(GOLD["query_words"], ids, GOLD["N"])
    assert np.floor(lik.astype(np.float64) * 1000).astype(int).tolist() == GOLD["expected_floor_likelihood_x1000"]


def build_map_pair(n_words, n_sigs, feats, id_stride=1):
    eng, o, vocab, ids = make_pair(n_words, id_stride=id_stride)
    m = synth.make_map(ids, n_sigs, feats, seed=2)
    eng.load_csr(m.word_ids, m.row_ptr, m.sig, m.cnt)
    o.load_csr(m.word_ids, m.row_ptr, m.sig, m.cnt)
    return eng, o, vocab, ids, m


def test_score_matches_oracle():
    eng, o, vocab, ids, m = build_map_pair(6000, 400, 300, id_stride=2)
    rng = np.random.default_rng(12)
    for k in range(5):
        qwords = m.sig_words[rng.integers(0, 400)].copy()
        qwords[::9] = 0
        qwords[1::13] = -7
        sig_ids = np.concatenate([m.sig_ids[rng.permutation(400)[:350]], [5000]]).astype(np.int32)
        got = eng.score(qwords, sig_ids, 401)
        want = o.likelihood(qwords, sig_ids, 401)
        assert_likelihood_close(got, want)
        assert got[-1] == 0.0


def test_refs_added_incrementally_equal_bulk_load():
    eng, o, vocab, ids, m = build_map_pair(2000, 60, 120)
    eng2 = Engine()
    eng2.add_words(ids, vocab)
    eng2.update()
    for s in range(60):
        eng2.add_refs(int(m.sig_ids[s]), m.sig_words[s])
    assert eng2.total_refs() == eng.total_refs() == 60 * 120
    for w in m.word_ids[::37]:
        a = eng.get_refs(int(w)); b = eng2.get_refs(int(w)); c = o.get_refs(int(w))
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])
    q = m.sig_words[3]
    assert np.array_equal(eng.score(q, m.sig_ids, 60), eng2.score(q, m.sig_ids, 60))  # fixed-point sums: order independent


# ---------------------------------------------------------------- fused batch -------------------
@pytest.mark.parametrize("incremental", [True, False])
def test_localize_batch_matches_oracle(incremental):
    eng, o, vocab, ids, m = build_map_pair(8000, 500, 400)
    o.set_params(incremental, 0.8, True)
    B, F = 6, 400
    q, places = synth.make_query_frames(vocab, ids, m, B, F, seed=3)
    words, like = eng.localize_batch(q, B, m.sig_ids, 501, incremental=incremental)
    for b in range(B):
        w_o, l_o = o.localize(q[b * F:(b + 1) * F], 9999, m.sig_ids, 501)
        assert np.array_equal(words[b], w_o), f"frame {b}"
        assert_likelihood_close(like[b], l_o)
        assert int(m.sig_ids[np.argmax(like[b])]) == int(places[b])  # the revisited place wins
    # the engine was not mutated
    assert eng.size() == 8000 and eng.not_indexed_size() == 0 and eng.total_refs() == 500 * 400


def test_batch_equals_single_frames():
    eng, o, vocab, ids, m = build_map_pair(5000, 300, 256)
    q, _ = synth.make_query_frames(vocab, ids, m, 5, 256, seed=4)
    words, like = eng.localize_batch(q, 5, m.sig_ids, 301)
    for b in range(5):
        w1, l1 = eng.localize_batch(q[b * 256:(b + 1) * 256], 1, m.sig_ids, 301)
        assert np.array_equal(w1[0], words[b]) and np.array_equal(l1[0], like[b])


def test_vwdictionary_mirror_interface():
    d = VWDictionaryB200({"Kp/NndrRatio": "0.8", "Kp/IncrementalDictionary": "true"})
    o = orc.OracleDictionary()
    rng = np.random.default_rng(1)
    assert d.addNewWords(np.zeros((0, 32), np.uint8), 1) == []
    for t in range(1, 5):
        f = rng.integers(0, 256, (120, 32), dtype=np.uint8)
        f[60:] = synth.flip_bits(f[:60], 0.03, rng)
        d.update(); o.update()
        assert d.addNewWords(f, t) == o.add_new_words(f, t).tolist()
    assert d.getLastWordId() == o.last_word_id
    assert not d.addWordRef(10 ** 6, 1)
    some = d.getLastWordId()
    assert d.addWordRef(some, 9) and d.getReferences(some).get(9) == 1
    d.update()
    lik = d.computeLikelihood([some, some, 0, -1], [1, 2, 3, 4, 9], 5)
    assert set(lik) == {1, 2, 3, 4, 9} and lik[9] > 0


# ---------------------------------------------------------------- adjustLikelihood ---------------
@pytest.mark.parametrize("ratio", [0, 1])
def test_adjust_likelihood_bit_exact(ratio):
    """Rtabmap::adjustLikelihood: mean / standard deviation of the values > 0 summed in list order, so the CUDA result has to
    equal the restated float arithmetic bit for bit; rows with no / one / equal positive values take the reference's branches."""
    rng = np.random.default_rng(40 + ratio)
    eng = Engine()
    rows = []
    for n in (1, 2, 7, 300, 2048, 2049, 10000):
        r = np.abs(rng.standard_normal(n)).astype(np.float32) * 0.01
        r[rng.random(n) < 0.6] = 0.0          # most signatures share no word with the frame
        if n >= 7:
            r[rng.integers(0, n, 3)] = rng.random(3).astype(np.float32)   # a few loop-closure candidates
        rows.append(r)
    rows.append(np.zeros(50, np.float32))                       # nothing in common with any signature
    one = np.zeros(50, np.float32); one[17] = 0.25; rows.append(one)   # a single positive value: variance 0
    rows.append(np.full(64, 0.125, np.float32))                 # all equal: stdDev 0
    for r in rows:
        got = eng.adjust_likelihood(r, ratio)
        want = orc.adjust_likelihood(np.concatenate([[0.0], r]).astype(np.float32), ratio)
        assert got.shape == want.shape and np.array_equal(got, want), (len(r), ratio)
    batch = np.stack([rows[3], rows[3][::-1].copy(), np.zeros(300, np.float32)])
    got = eng.adjust_likelihood(batch, ratio)
    for b in range(3):
        assert np.array_equal(got[b], orc.adjust_likelihood(np.concatenate([[0.0], batch[b]]).astype(np.float32), ratio))
