"""Float descriptors (LCD_DESC_F32: SURF-64 / SURF-128 / SIFT sized rows) through the C ABI against the oracle, whose squared-L2
2-NN is pinned to the reference's own compiled rtflann (tests/test_oracle_ref.py).  Bar: identical neighbours, bit-identical
distances, identical word ids for the incremental / fixed quantiser and findNN, likelihood within 1e-4."""
import numpy as np
import pytest

from oracle import oracle_py as orc
from rtabmap_b200 import Engine, synth
from rtabmap_b200.capi import LCD_DESC_F32

pytestmark = pytest.mark.gpu


def float_vocab(n, dim, seed):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, dim)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)          # SURF / SIFT descriptors are unit length
    return v.astype(np.float32)


def pair(n_words, dim, seed=1, incremental=True, nndr=0.8):
    vocab = float_vocab(n_words, dim, seed)
    ids = np.arange(1, n_words + 1, dtype=np.int32) * 2
    eng = Engine(desc_type=LCD_DESC_F32, desc_dim=dim)
    o = orc.OracleDictionary(1, dim, incremental, nndr, True)
    if n_words:
        eng.add_words(ids, vocab)
        o.add_words(ids, vocab)
    last = int(ids.max()) if n_words else 0
    eng.last_word_id = last
    o.last_word_id = last
    eng.update()
    o.update()
    return eng, o, vocab, ids


def noisy(rows, sigma, rng):
    q = rows + sigma * rng.standard_normal(rows.shape).astype(np.float32)
    return (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("dim,n_words", [(64, 0), (64, 1), (64, 2), (64, 700), (64, 5000), (128, 900)])
def test_knn2_float_bit_exact(dim, n_words):
    eng, o, vocab, ids = pair(n_words, dim, seed=dim + n_words)
    rng = np.random.default_rng(7)
    q = float_vocab(300, dim, 99)
    if n_words >= 700:
        q[:100] = noisy(vocab[rng.integers(0, n_words, 100)], 0.05, rng)
        q[100] = vocab[5]
        q[101] = vocab[5]
    g = eng.knn2(q)
    w = o.knn2(q)
    assert np.array_equal(g[0], w[0]) and np.array_equal(g[2], w[2])
    assert np.array_equal(g[1].view(np.uint32), w[1].view(np.uint32)) and np.array_equal(g[3].view(np.uint32), w[3].view(np.uint32))


def test_knn2_float_ties_resolve_to_lowest_row():
    vocab = float_vocab(500, 64, 3)
    vocab[77] = vocab[9]
    vocab[300] = vocab[9]
    ids = np.arange(1, 501, dtype=np.int32)
    eng = Engine(desc_type=LCD_DESC_F32, desc_dim=64)
    eng.add_words(ids, vocab)
    eng.update()
    id1, d1, id2, d2 = eng.knn2(vocab[9:10])
    assert (id1[0], d1[0], id2[0], d2[0]) == (10, 0.0, 78, 0.0)


@pytest.mark.parametrize("dim", [64, 128])
def test_incremental_stream_float(dim):
    """Mapping mode: frames quantised one after the other with update() in between; words created inside a frame are compared
    together (Kp/NewWordsComparedTogether) — ids, number of new words and the index must follow the oracle frame by frame."""
    eng, o, vocab, ids = pair(1500, dim, seed=11)
    rng = np.random.default_rng(13)
    for t in range(1, 6):
        f = float_vocab(160, dim, 100 + t)
        f[:60] = noisy(vocab[rng.integers(0, 1500, 60)], 0.03, rng)      # revisits of known words
        f[100:130] = noisy(f[60:90], 0.02, rng)                         # near-duplicates inside the frame
        prev_last = o.last_word_id
        got, n_new = eng.quantize(f, t)
        want = o.add_new_words(f, t)
        assert np.array_equal(got, want), f"frame {t}"
        assert n_new == len(np.unique(want[want > prev_last])) and n_new > 0
        assert eng.last_word_id == o.last_word_id and eng.not_indexed_size() == o.not_indexed_size()
        eng.update()
        o.update()
        assert eng.indexed_size() == o.indexed_size()
    probe = noisy(vocab[rng.integers(0, 1500, 50)], 0.04, rng)
    assert np.array_equal(eng.find_nn(probe), o.find_nn(probe))


def test_fixed_dictionary_and_localize_float():
    eng, o, vocab, ids = pair(3000, 64, seed=21, incremental=False)
    m = synth.make_map(ids, 200, 150, seed=2)
    eng.load_csr(m.word_ids, m.row_ptr, m.sig, m.cnt)
    o.load_csr(m.word_ids, m.row_ptr, m.sig, m.cnt)
    rng = np.random.default_rng(5)
    B, F = 3, 150
    q = np.concatenate([noisy(vocab[rng.integers(0, 3000, F)], 0.05, rng) for _ in range(B)])
    words, like = eng.localize_batch(q, B, m.sig_ids, 201, incremental=False)
    for b in range(B):
        w_o, l_o = o.localize(q[b * F:(b + 1) * F], 9999, m.sig_ids, 201)
        assert np.array_equal(words[b], w_o), f"frame {b}"
        assert np.allclose(like[b], l_o, atol=1e-4, rtol=1e-4)


def test_float_engine_rejects_binary_only_calls():
    eng = Engine(desc_type=LCD_DESC_F32, desc_dim=64)
    with pytest.raises(Exception):
        eng.match_pairs(np.zeros((1, 8, 64), np.float32), np.zeros((1, 8, 64), np.float32), [8], [8])
