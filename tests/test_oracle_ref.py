"""Pin the restated linear 2-NN scan (oracle/oracle.cpp) against the REFERENCE'S OWN rtflann,
compiled from /root/reference by oracle/Makefile into oracle/_ref/libref_flann.so.
Covers Hamming (ORB) and squared-L2 (SURF) incl. exact ties and fewer than two rows."""
import numpy as np
import pytest

from oracle import oracle_py as orc

pytestmark = pytest.mark.skipif(orc.ref_lib() is None, reason="oracle/_ref not built (no /root/reference)")


@pytest.mark.parametrize("rows,dim", [(1, 32), (2, 32), (3, 32), (777, 32), (4096, 32), (1000, 16), (500, 64)])
def test_hamming_matches_rtflann(rows, dim):
    rng = np.random.default_rng(rows * 7 + dim)
    data = rng.integers(0, 256, (rows, dim), dtype=np.uint8)
    q = rng.integers(0, 256, (64, dim), dtype=np.uint8)
    if rows > 10:
        data[5] = data[1]
        data[9] = data[1]          # three identical rows -> ties resolved to the lowest row
        q[0] = data[1]
        q[1] = data[rows - 1]
        q[2] = np.bitwise_xor(data[3], 1)
    i_ref, d_ref = orc.ref_knn2(data, q)
    i_orc, d_orc = orc.knn2_raw(data, q)
    assert np.array_equal(i_ref, i_orc)
    valid = i_ref >= 0
    assert np.array_equal(d_ref[valid], d_orc[valid])


@pytest.mark.parametrize("rows,dim", [(2, 64), (1500, 64), (800, 128)])
def test_l2_matches_rtflann_bit_exact(rows, dim):
    rng = np.random.default_rng(rows + dim)
    data = rng.standard_normal((rows, dim)).astype(np.float32)
    data /= np.linalg.norm(data, axis=1, keepdims=True)
    q = (data[rng.integers(0, rows, 48)] + 0.02 * rng.standard_normal((48, dim))).astype(np.float32)
    if rows > 10:
        data[7] = data[2]
        q[0] = data[2]
    i_ref, d_ref = orc.ref_knn2(data, q)
    i_orc, d_orc = orc.knn2_raw(data, q)
    assert np.array_equal(i_ref, i_orc)
    # same float summation order as rtflann::L2 (dist.h:158-166): identical bits
    assert np.array_equal(d_ref.view(np.uint32), d_orc.view(np.uint32))
