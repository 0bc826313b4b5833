"""Cold start of the engine from an RTAB-Map database, and saving the dictionary back.

Host-side I/O only: SQLite rows in, the arrays the C ABI takes out.  Follows the reference's SQLite driver
(corelib/src/DBDriverSqlite3.cpp) on the tables of corelib/src/resources/DatabaseSchema.sql.in:

  Word(id, descriptor_size, descriptor BLOB, time_enter)                      :66-72
  Feature(node_id, word_id, pos_x, pos_y, size, dir, response, octave, depth_x, depth_y, depth_z, descriptor_size, descriptor)   :74-89

  read_dictionary      DBDriverSqlite3::loadQuery(VWDictionary &, lastStateOnly)  :3541-3620: words ORDER BY id; a blob as long as
                       descriptor_size is CV_8U, one 4x as long is CV_32F, anything else is fatal; last word id = the largest id
  read_inverted_index  what Memory::enableWordsRef (Memory.cpp:7005-7030) rebuilds from the signatures' word lists: one reference per
                       Feature row with word_id > 0 (multiplicity kept), and Ni = ALL Feature rows of the node
                       (DBDriverSqlite3::getInvertedIndexNiQuery :2775-2790: count(word_id) ... WHERE node_id = ?)
  write_dictionary     DBDriverSqlite3::saveQuery(const std::list<VisualWord *> &) :4734: INSERT INTO Word(id, descriptor_size, descriptor)

The FLANN index blob (Admin.dictionary_index, :3622-3660) is not read: the engine searches exhaustively and keeps no tree.
"""
from __future__ import annotations

import sqlite3
from dataclasses import dataclass

import numpy as np


class DbFormatError(ValueError):
    pass


@dataclass
class DictionaryRows:
    ids: np.ndarray          # int32 [n], ascending
    descriptors: np.ndarray  # uint8 [n, size] or float32 [n, size]
    last_word_id: int


@dataclass
class InvertedIndexRows:
    word_ids: np.ndarray  # int32 [n_words], ascending: words that have at least one reference
    row_ptr: np.ndarray   # int64 [n_words + 1]
    sig: np.ndarray       # int32 [nnz] signature (node) id, ascending inside a word
    cnt: np.ndarray       # int32 [nnz] occurrences of the word in that signature
    sig_ids: np.ndarray   # int32 [n_sigs], ascending: every node that has Feature rows
    ni: np.ndarray        # int32 [n_sigs] number of Feature rows of the node (Memory::getNi)


def _connect(path: str) -> sqlite3.Connection:
    con = sqlite3.connect(f"file:{path}?mode=ro", uri=True)
    con.row_factory = None
    return con


def read_dictionary(path: str, last_state_only: bool = False) -> DictionaryRows:
    con = _connect(path)
    try:
        q = "SELECT id, descriptor_size, descriptor FROM Word "
        if last_state_only:
            q += "WHERE time_enter >= (SELECT MAX(time_enter) FROM Info) "  # databases >= 0.11.11 (:3557-3561)
        rows = con.execute(q + "ORDER BY id;").fetchall()
    finally:
        con.close()
    if not rows:
        return DictionaryRows(np.zeros(0, np.int32), np.zeros((0, 0), np.uint8), 0)
    size = int(rows[0][1])
    first = len(rows[0][2])
    if first == size:
        dtype, width = np.uint8, size
    elif first // 4 == size:
        dtype, width = np.float32, size
    else:
        raise DbFormatError(f"Saved buffer size ({first} bytes) is not the same as descriptor size ({size})")
    ids = np.empty(len(rows), np.int32)
    desc = np.empty((len(rows), width), dtype)
    row_bytes = width * np.dtype(dtype).itemsize
    for i, (wid, dsize, blob) in enumerate(rows):
        # one dictionary = one descriptor type and width (VWDictionary::update asserts it when it builds the index, VWDictionary.cpp:557-558, :600-601, :670-671)
        if int(dsize) != size or len(blob) != row_bytes:
            raise DbFormatError(f"word {wid}: descriptor of {len(blob)} bytes / size {dsize} in a dictionary of {row_bytes} bytes / size {size}")
        ids[i] = wid
        desc[i] = np.frombuffer(blob, dtype=dtype, count=width)
    return DictionaryRows(ids, desc, int(ids.max()))


def read_inverted_index(path: str, node_ids=None) -> InvertedIndexRows:
    """node_ids: restrict to these signatures (the working memory being reloaded); None = every node of the Feature table."""
    con = _connect(path)
    try:
        rows = np.array(con.execute("SELECT node_id, word_id FROM Feature;").fetchall(), dtype=np.int64).reshape(-1, 2)
    finally:
        con.close()
    if node_ids is not None:
        rows = rows[np.isin(rows[:, 0], np.asarray(list(node_ids), np.int64))]
    if rows.shape[0] == 0:
        z = np.zeros(0, np.int32)
        return InvertedIndexRows(z, np.zeros(1, np.int64), z, z, z, z)
    sig_ids, ni = np.unique(rows[:, 0], return_counts=True)
    refs = rows[rows[:, 1] > 0]
    # (word, signature) pairs with their multiplicity, words ascending, signatures ascending inside a word
    key = refs[:, 1] * (int(sig_ids.max()) + 1) + refs[:, 0]
    uniq, cnt = np.unique(key, return_counts=True)
    w = uniq // (int(sig_ids.max()) + 1)
    s = uniq % (int(sig_ids.max()) + 1)
    word_ids, first = np.unique(w, return_index=True)
    row_ptr = np.concatenate([first, [len(w)]]).astype(np.int64)
    return InvertedIndexRows(word_ids.astype(np.int32), row_ptr, s.astype(np.int32), cnt.astype(np.int32), sig_ids.astype(np.int32), ni.astype(np.int32))


def write_dictionary(path: str, ids, descriptors) -> int:
    """Append words to the Word table (the table must exist: the reference creates it from DatabaseSchema.sql); returns the row count."""
    ids = np.asarray(ids, np.int32)
    d = np.ascontiguousarray(descriptors)
    if d.dtype not in (np.uint8, np.float32) or d.ndim != 2 or len(d) != len(ids):
        raise DbFormatError("descriptors must be uint8 or float32 [n, size], one row per id")
    con = sqlite3.connect(path)
    try:
        with con:
            con.executemany("INSERT INTO Word(id, descriptor_size, descriptor) VALUES(?,?,?);",
                            [(int(i), int(d.shape[1]), d[k].tobytes()) for k, i in enumerate(ids)])
    finally:
        con.close()
    return len(ids)


def cold_start(engine, path: str, node_ids=None) -> tuple[DictionaryRows, InvertedIndexRows]:
    """Load dictionary + inverted index of a database into an engine created for its descriptor type: the device-side equivalent of
    Memory::init -> DBDriver::load(VWDictionary) + enableWordsRef for the reloaded signatures."""
    words = read_dictionary(path)
    index = read_inverted_index(path, node_ids)
    if len(words.ids):
        engine.add_words(words.ids, words.descriptors)
        engine.last_word_id = words.last_word_id
        engine.update()
    if len(index.word_ids):
        missing = np.setdiff1d(index.word_ids, words.ids)
        if len(missing):
            raise DbFormatError(f"{len(missing)} words referenced by Feature rows are not in the Word table (first: {int(missing[0])})")
        engine.load_csr(index.word_ids, index.row_ptr, index.sig, index.cnt)
    if len(index.sig_ids):
        engine.set_ni(index.sig_ids, index.ni)
    return words, index
