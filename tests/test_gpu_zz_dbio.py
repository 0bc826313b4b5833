"""Cold start of the engine from an RTAB-Map database: an engine loaded through rtabmap_b200.dbio from the Word /
Feature tables must answer exactly like an engine given the same dictionary and inverted index directly."""
import numpy as np
import pytest

from rtabmap_b200 import Engine, dbio, synth
from test_dbio import make_db

pytestmark = pytest.mark.gpu


def test_cold_start_from_a_database_equals_direct_loading(tmp_path):
    vocab = synth.make_binary_vocabulary(2048, 32, 1)
    ids = np.arange(1, 2049, dtype=np.int32)
    smap = synth.make_map(ids, 120, 150, seed=2)
    q, places = synth.make_query_frames(vocab, ids, smap, 3, 150, seed=3)
    db = str(tmp_path / "map.db")
    # one Feature row per word occurrence of every signature (Signature::getWords), as DBDriverSqlite3 stores them
    feats = [(int(s), int(w)) for s, row in zip(smap.sig_ids, smap.sig_words) for w in row]
    make_db(db, ids, vocab, feats)

    direct = Engine()
    direct.add_words(ids, vocab)
    direct.last_word_id = 2048
    direct.update()
    direct.load_csr(smap.word_ids, smap.row_ptr, smap.sig, smap.cnt)
    direct.set_ni(smap.sig_ids, smap.ni)

    cold = Engine()
    words, index = dbio.cold_start(cold, db)
    assert np.array_equal(words.ids, ids) and np.array_equal(index.word_ids, smap.word_ids) and np.array_equal(index.row_ptr, smap.row_ptr)
    assert np.array_equal(index.sig, smap.sig) and np.array_equal(index.cnt, smap.cnt) and np.array_equal(index.ni, smap.ni)

    w0, l0 = direct.localize_batch(q, 3, smap.sig_ids, 121)
    w1, l1 = cold.localize_batch(q, 3, smap.sig_ids, 121)
    assert np.array_equal(w0, w1) and np.array_equal(l0, l1)
    assert l1.max() > 0
