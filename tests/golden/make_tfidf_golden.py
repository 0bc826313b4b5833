"""Make tests/golden/tfidf_golden.json from the reference's only known-answer test for the
scoring stage: archive/2010-LoopClosure/Tests/TestComputeLikelihood.m with its two fixtures
(090306-3_db-Signatures.txt, 090306-3_db-Dictionary.txt).

The MATLAB test first refreshes the virtual place (Bayes/updateCommonSignature.m +
updateDictionary.m), then asserts floor(computeLikelihood(...)*1000) against an 83-vector.
This script ports that pre-step, checks the test's own 'sign' assertion, and writes the
derived inverted index (word -> [(signature, count)]), ni per signature, the query's words and
the expected vector (copied from the .m file) so that the oracle and the CUDA path can be
pinned without /root/reference at test time.

Run here (needs /root/reference):  python tests/golden/make_tfidf_golden.py
"""
import json
import math
import re
from pathlib import Path

import numpy as np

REF = Path("/root/reference/archive/2010-LoopClosure")
OUT = Path(__file__).resolve().parent / "tfidf_golden.json"


def dlmread(path):
    rows = []
    for line in path.read_text().splitlines()[1:]:
        vals = [int(v) for v in line.split()]
        if vals:
            rows.append(vals)
    width = max(len(r) for r in rows)
    m = np.zeros((len(rows), width), dtype=np.int64)
    for i, r in enumerate(rows):
        m[i, : len(r)] = r
    return m


def update_dictionary(D, sign):
    sign_id = sign[0]
    words = [w for w in sign[1:] if w != 0]
    for w in words:
        idx = np.where(D[:, 0] == w)[0]
        if len(idx) == 0:
            D = np.vstack([D, np.zeros((1, D.shape[1]), dtype=D.dtype)])
            D[-1, 0] = w
            D[-1, 1] = sign_id
        else:
            r = idx[0]
            zeros = np.where(D[r, :] == 0)[0]
            if len(zeros) == 0:
                D = np.hstack([D, np.zeros((D.shape[0], 1), dtype=D.dtype)])
                D[r, -1] = sign_id
            else:
                D[r, zeros[0]] = sign_id
    return D


def update_common_signature(M, D):
    cs = list(M[0, :])
    cs_id = M[0, 0]
    for w in cs[1:]:
        idx = np.where(D[:, 0] == w)[0]
        if len(idx):
            r = idx[0]
            D[r, D[r, :] == cs_id] = 0
    cs = [cs_id]
    mem_size = M.shape[0] - 1
    nb = 0
    if mem_size > 0:
        nb = int(np.count_nonzero(D[:, 1:])) // mem_size
    if nb > 0:
        counts = np.count_nonzero(D[:, 1:], axis=1)
        lst = sorted(zip(counts.tolist(), D[:, 0].tolist()))
        added = 0
        n = len(lst)
        for i in range(n - 1, -1, -1):
            if i != n - 1 and len(cs) > 1:
                ratio = lst[i + 1][0] // lst[i][0] if lst[i][0] else 10 ** 9
                ln = len(cs)
                for _ in range(2, ratio + 1):
                    for k in range(1, ln):
                        cs.append(cs[k])
                        added += 1
                        if added >= nb:
                            break
                    if added >= nb:
                        break
            if added < nb:
                cs.append(lst[i][1])
                added += 1
            if added >= nb:
                break
        cs = cs + [0] * (M.shape[1] - len(cs))
        D = update_dictionary(D, cs)
    return cs, D


def main():
    M = dlmread(REF / "Tests/090306-3_db-Signatures.txt")
    D = dlmread(REF / "Tests/090306-3_db-Dictionary.txt")
    cs, D = update_common_signature(M, D)
    if len(cs) > M.shape[1]:
        M = np.hstack([M, np.zeros((M.shape[0], len(cs) - M.shape[1]), dtype=M.dtype)])
    M[0, : len(cs)] = cs
    M[0, len(cs):] = 0

    src = (REF / "Tests/TestComputeLikelihood.m").read_text()
    vecs = re.findall(r"\[([0-9,\s;]+)\]", src)
    sign_expected = [int(v) for v in vecs[0].replace(";", "").split(",") if v.strip()]
    lik_expected = [int(v) for v in vecs[1].replace(";", "").split(",") if v.strip()]
    sign = M[-1, :].tolist()
    assert sign[: len(sign_expected)] == sign_expected[: len(sign)] or sign == sign_expected[: len(sign)], "sign is not valid!"

    # derived structures
    N = M.shape[0]
    sig_ids = M[:, 0].tolist()
    ni = [int(np.count_nonzero(M[r, 1:] > 0)) for r in range(N)]
    words = {}
    for r in range(D.shape[0]):
        w = int(D[r, 0])
        refs = sorted(set(int(v) for v in D[r, 1:] if v != 0))
        entry = []
        for s in refs:
            row = np.where(M[:, 0] == s)[0][0]
            nwi = int(np.count_nonzero(M[row, 1:] == w))
            entry.append([s, nwi])
        words[w] = entry
    query_words = [int(v) for v in sign[1:] if v != 0]

    # double-precision replay of Bayes/computeLikelihood.m as a self-check of this port
    lik = np.zeros(N)
    for w in sorted(set(query_words)):
        refs = words[w]
        nw = len(refs)
        logn = math.log10(N / nw)
        if logn != 0:
            for s, nwi in refs:
                pos = sig_ids.index(s)
                nwi_m = int(np.count_nonzero(M[pos, 1:] == w))
                if ni[pos]:
                    lik[pos] += (nwi_m * logn) / ni[pos]
    got = np.floor(lik * 1000).astype(int).tolist()
    assert got == lik_expected, (got, lik_expected)

    OUT.write_text(json.dumps({
        "source": "archive/2010-LoopClosure/Tests/TestComputeLikelihood.m (+ fixtures, after updateCommonSignature)",
        "N": N, "sig_ids": sig_ids, "ni": ni, "query_sig": int(sign[0]), "query_words": query_words,
        "words": {str(k): v for k, v in words.items()},
        "expected_floor_likelihood_x1000": lik_expected,
    }))
    print("wrote", OUT, "N=", N, "words=", len(words), "self-check OK")


if __name__ == "__main__":
    main()
