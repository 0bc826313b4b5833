"""Host-side logic of word-range sharding across GPUs (BASELINE config 5).

Rank r of G owns a contiguous range of vocabulary rows (rows are in ascending word id, the
search order of VWDictionary::update) and the posting lists of exactly those words.  Per batch:
  1. every rank finds its local top-2 of each descriptor as packed keys (dist << 22 | GLOBAL row);
  2. all-gather of the keys; every rank merges them (unsigned min-2 = the reference's
     (distance, lowest row) order) and runs the replicated NNDR / new-word pass;
  3. every rank scores the words it owns into exact fixed-point int64 sums; all-reduce(sum).
No computation of the hot path happens here: this module only partitions inputs and packs keys
(the numpy merge below is the specification the gloo CPU test checks the partitioning against).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

KEY_SHIFT = 22
KEY_ROW_MASK = (1 << KEY_SHIFT) - 1
KEY_NONE = np.uint32(0xFFFFFFFF)


def shard_rows(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row range [r0, r1) of `rank`."""
    base, rem = divmod(n_rows, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def shard_csr(word_ids: np.ndarray, row_ptr: np.ndarray, sig: np.ndarray, cnt: np.ndarray, owned_ids: np.ndarray):
    """Subset of a word-major CSR inverted index restricted to the words in `owned_ids`."""
    owned = np.isin(word_ids, owned_ids)
    lens = np.diff(row_ptr)
    keep = np.repeat(owned, lens)
    new_ptr = np.concatenate([[0], np.cumsum(lens[owned])]).astype(np.int64)
    return word_ids[owned], new_ptr, sig[keep], cnt[keep]


def pack_keys(dist: np.ndarray, row: np.ndarray) -> np.ndarray:
    return (dist.astype(np.uint32) << np.uint32(KEY_SHIFT)) | row.astype(np.uint32)


def merge_top2(keys: np.ndarray) -> np.ndarray:
    """keys [G, nq, 2] uint32 -> [nq, 2]: the two smallest keys per query (KEY_NONE = absent)."""
    g, nq, _ = keys.shape
    flat = np.sort(np.transpose(keys, (1, 0, 2)).reshape(nq, g * 2), axis=1)
    return flat[:, :2]
