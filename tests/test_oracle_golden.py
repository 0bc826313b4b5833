"""The oracle against the reference's golden vector for the scoring stage
(archive/2010-LoopClosure/Tests/TestComputeLikelihood.m, fixture made by
tests/golden/make_tfidf_golden.py), plus hand-checked cases of the sequential
addNewWords semantics (VWDictionary.cpp:1088-1219)."""
import numpy as np

from oracle import oracle_py as orc

from golden_util import GOLD, load_golden_into


def test_tfidf_golden_vector():
    remap = lambda s: 1000 if s == -1 else s  # the virtual place has id -1 in the 2010 fixtures
    d = orc.OracleDictionary()
    load_golden_into(d, remap)
    ids = [remap(s) for s in GOLD["sig_ids"]]
    lik = d.likelihood(GOLD["query_words"], ids, GOLD["N"])
    got = np.floor(lik.astype(np.float64) * 1000).astype(int).tolist()
    assert got == GOLD["expected_floor_likelihood_x1000"]


def _bits(*ones, n=32):
    v = np.zeros(n * 8, np.uint8)
    for o in ones:
        v[o] = 1
    return np.packbits(v)


def test_nndr_and_new_word_order():
    d = orc.OracleDictionary(nndr=0.8)
    base = np.stack([_bits(), _bits(*range(100, 160))])  # word 1 = zeros, word 2 = 60 bits set
    d.add_words([1, 2], base)
    d.last_word_id = 2
    d.update()
    q = np.stack([
        _bits(0, 1),                 # d(w1)=2, d(w2)=62   -> 2 <= 0.8*62 -> word 1
        _bits(*range(200, 256)),     # far from both (56 / 116): 56 > 0.8*116? no -> 56<=92.8 -> word 1
        _bits(*range(100, 130)),     # d(w1)=30, d(w2)=30  -> tie, 30 > 24 -> rejected -> new word 3
        _bits(*range(100, 131)),     # d(new3)=1, d(w1)=31, d(w2)=29 -> best new3 (1 <= 0.8*29) -> word 3
    ])
    ids = d.add_new_words(q, 7)
    assert ids.tolist() == [1, 1, 3, 3]
    assert d.not_indexed_size() == 1 and d.last_word_id == 3
    s, c = d.get_refs(3)
    assert s.tolist() == [7] and c.tolist() == [2]
    s, c = d.get_refs(1)
    assert s.tolist() == [7] and c.tolist() == [2]


def test_index_hits_win_ties_against_new_words():
    # one indexed word only: the first descriptor has < 2 results -> new word 2 (VWDictionary.cpp:1176-1179)
    d = orc.OracleDictionary(nndr=1.0)
    d.add_words([1], np.stack([_bits(*range(0, 10))]))
    d.last_word_id = 1
    d.update()
    q = np.stack([
        _bits(*range(40, 50)),                  # -> new word 2
        _bits(*range(0, 10), *range(40, 50)),   # d(w1)=10 == d(new2)=10: index hit was inserted first -> word 1
        _bits(*range(0, 9), *range(40, 50)),    # d(w1)=11, d(new2)=9 -> word 2
    ])
    ids = d.add_new_words(q, 1)
    assert ids.tolist() == [2, 1, 2]


def test_fixed_dictionary_takes_nearest():
    d = orc.OracleDictionary(incremental=False)
    d.add_words([5, 9], np.stack([_bits(), _bits(*range(0, 128))]))
    d.update()
    ids = d.add_new_words(np.stack([_bits(*range(0, 64)), _bits(*range(0, 65))]), 3)
    assert ids.tolist() == [5, 9]  # 64/64 tie -> lowest row (word 5); 65/63 -> word 9


def test_adjust_likelihood_matches_formula():
    lik = np.array([0.0, 0.1, 0.2, 0.9, 0.0, 0.15], np.float32)
    out = orc.adjust_likelihood(lik)
    vals = lik[1:][lik[1:] > 0]
    mean = np.float32(vals.mean())
    std = np.float32(vals.std(ddof=1))
    assert out[3] > 1.0 and np.isclose(out[3], (0.9 - (std - 1e-4)) / mean, rtol=1e-5)
    assert out[1] == 1.0 and out[4] == 1.0
    assert np.isclose(out[0], mean / std + 1.0, rtol=1e-5)


def test_read_only_localisation_equals_insert_and_roll_back():
    from rtabmap_b200 import synth

    vocab = synth.make_binary_vocabulary(3000, 32, 1)
    ids = np.arange(1, 3001, dtype=np.int32)
    m = synth.make_map(ids, 200, 150)
    d = orc.OracleDictionary()
    d.add_words(ids, vocab)
    d.last_word_id = 3000
    d.update()
    d.load_csr(m.word_ids, m.row_ptr, m.sig, m.cnt)
    q, _ = synth.make_query_frames(vocab, ids, m, 3, 150)
    for b in range(3):
        f = q[b * 150:(b + 1) * 150]
        w1, l1 = d.localize(f, 777, m.sig_ids, 201)
        w2, l2 = d.localize_ro(f, m.sig_ids, 201)
        assert np.array_equal(w1, w2) and np.array_equal(l1, l2)
        assert d.size() == 3000 and d.last_word_id == 3000 and d.total_refs() == 200 * 150
