"""Cold start from an RTAB-Map database, the part that needs no GPU: Word / Feature rows of a database with the
reference's table layout (corelib/src/resources/DatabaseSchema.sql.in:66-89) -> the arrays the C ABI takes, checked against a plain
dictionary-of-dictionaries restatement of what Memory::enableWordsRef / getNi would rebuild."""
import sqlite3

import numpy as np
import pytest

from rtabmap_b200 import dbio

# the two tables of the reference schema this module reads, column for column (names and order are the format)
SCHEMA = """
CREATE TABLE Word (id INTEGER NOT NULL, descriptor_size INTEGER NOT NULL, descriptor BLOB NOT NULL, time_enter DATE, PRIMARY KEY (id));
CREATE TABLE Feature (node_id INTEGER NOT NULL, word_id INTEGER NOT NULL, pos_x FLOAT NOT NULL, pos_y FLOAT NOT NULL, size INTEGER NOT NULL,
                      dir FLOAT NOT NULL, response FLOAT NOT NULL, octave INTEGER NOT NULL, depth_x FLOAT, depth_y FLOAT, depth_z FLOAT,
                      descriptor_size INTEGER, descriptor BLOB);
"""


def make_db(path, words, desc, features):
    con = sqlite3.connect(path)
    con.executescript(SCHEMA)
    con.executemany("INSERT INTO Word(id, descriptor_size, descriptor) VALUES(?,?,?);", [(int(i), desc.shape[1], desc[k].tobytes()) for k, i in enumerate(words)])
    con.executemany("INSERT INTO Feature(node_id, word_id, pos_x, pos_y, size, dir, response, octave) VALUES(?,?,0,0,31,0,0,0);", [(int(n), int(w)) for n, w in features])
    con.commit()
    con.close()


@pytest.mark.parametrize("dtype,width", [(np.uint8, 32), (np.float32, 64)])
def test_dictionary_round_trip(tmp_path, dtype, width):
    rng = np.random.default_rng(3)
    ids = np.sort(rng.choice(np.arange(1, 500), 120, replace=False)).astype(np.int32)
    desc = rng.integers(0, 256, (120, width)).astype(dtype) if dtype == np.uint8 else rng.normal(size=(120, width)).astype(np.float32)
    db = str(tmp_path / "map.db")
    order = rng.permutation(120)                      # rows are stored in any order, read ORDER BY id
    make_db(db, ids[order], desc[order], [])
    got = dbio.read_dictionary(db)
    assert got.descriptors.dtype == dtype and np.array_equal(got.ids, ids) and np.array_equal(got.descriptors, desc)
    assert got.last_word_id == int(ids.max())
    # saving back: new words appended by write_dictionary are read again
    more = np.array([600, 601], np.int32)
    dbio.write_dictionary(db, more, desc[:2])
    again = dbio.read_dictionary(db)
    assert np.array_equal(again.ids, np.concatenate([ids, more])) and np.array_equal(again.descriptors[-2:], desc[:2]) and again.last_word_id == 601


def test_inverted_index_matches_the_reference_bookkeeping(tmp_path):
    rng = np.random.default_rng(4)
    words = np.arange(1, 201, dtype=np.int32)
    desc = rng.integers(0, 256, (200, 32)).astype(np.uint8)
    feats = []
    for node in range(1, 31):
        k = int(rng.integers(20, 60))
        w = rng.integers(1, 201, k)
        w[rng.random(k) < 0.1] = -1                   # invalid words (no reference, but they count in Ni)
        w[rng.random(k) < 0.15] = w[0]                # repeated words: multiplicity is kept
        feats += [(node, int(x)) for x in w]
    db = str(tmp_path / "map.db")
    make_db(db, words, desc, feats)
    idx = dbio.read_inverted_index(db)
    # restatement: VisualWord::_references (word -> {signature: count}) built by addWordRef per valid key, Ni = words.size()
    refs, ni = {}, {}
    for node, w in feats:
        ni[node] = ni.get(node, 0) + 1
        if w > 0:
            refs.setdefault(w, {}).setdefault(node, 0)
            refs[w][node] += 1
    assert list(idx.sig_ids) == sorted(ni) and [int(x) for x in idx.ni] == [ni[n] for n in sorted(ni)]
    assert list(idx.word_ids) == sorted(refs)
    for k, w in enumerate(idx.word_ids):
        a, b = idx.row_ptr[k], idx.row_ptr[k + 1]
        assert dict(zip(idx.sig[a:b].tolist(), idx.cnt[a:b].tolist())) == refs[int(w)]
        assert np.all(np.diff(idx.sig[a:b]) > 0)
    # a working-memory subset
    sub = dbio.read_inverted_index(db, node_ids=[3, 7, 9])
    assert list(sub.sig_ids) == [3, 7, 9] and set(np.unique(sub.sig)) <= {3, 7, 9}
    assert int(sub.cnt.sum()) == sum(1 for n, w in feats if n in (3, 7, 9) and w > 0)


def test_a_blob_that_is_neither_bytes_nor_floats_is_fatal(tmp_path):
    db = str(tmp_path / "bad.db")
    con = sqlite3.connect(db)
    con.executescript(SCHEMA)
    con.execute("INSERT INTO Word(id, descriptor_size, descriptor) VALUES(1, 32, ?);", (bytes(48),))
    con.commit()
    con.close()
    with pytest.raises(dbio.DbFormatError):
        dbio.read_dictionary(db)


def test_empty_database(tmp_path):
    db = str(tmp_path / "empty.db")
    make_db(db, np.zeros(0, np.int32), np.zeros((0, 32), np.uint8), [])
    assert len(dbio.read_dictionary(db).ids) == 0
    idx = dbio.read_inverted_index(db)
    assert len(idx.word_ids) == 0 and list(idx.row_ptr) == [0]
