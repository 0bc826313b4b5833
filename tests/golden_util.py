"""Shared loader of tests/golden/tfidf_golden.json (made by tests/golden/make_tfidf_golden.py)."""
import json
from pathlib import Path

import numpy as np

GOLD = json.loads((Path(__file__).parent / "golden" / "tfidf_golden.json").read_text())


def load_golden_into(d, remap):
    """Feed the golden inverted index into a dictionary-like object (oracle or engine)."""
    word_ids = sorted(int(w) for w in GOLD["words"])
    desc = np.zeros((len(word_ids), 32), np.uint8)
    desc[:, :4] = np.asarray(word_ids, dtype=np.uint32).view(np.uint8).reshape(-1, 4)
    d.add_words(word_ids, desc)
    d.update()
    rp = [0]
    sig, cnt = [], []
    for w in word_ids:
        for s, c in GOLD["words"][str(w)]:
            sig.append(remap(s))
            cnt.append(c)
        rp.append(len(sig))
    d.load_csr(word_ids, rp, sig, cnt)
    d.set_ni([remap(s) for s in GOLD["sig_ids"]], GOLD["ni"])
