"""CPU: the C Chamfer oracle against outputs of the reference's own extension
(tests/golden/chamfer_ref_sm100a.npz, produced on a B200 by oracle/make_golden_chamfer.py from the
unmodified chamfer3D.cu): distances and indices bit-exact, gradients to atomic-add rounding."""
import os

import numpy as np
import pytest

import chamfer_cases as CC
from conftest import GOLDEN
from oracle import chamfer_oracle as CO

PATH = os.path.join(GOLDEN, "chamfer_ref_sm100a.npz")


@pytest.mark.skipif(not os.path.exists(PATH), reason="golden vectors not generated yet (oracle/make_golden_chamfer.py on a GPU box)")
@pytest.mark.parametrize("name", list(CC.GOLDEN_CASES))
def test_oracle_matches_reference_extension(name):
    gold = np.load(PATH)
    a, b, g1, g2 = CC.case(name)
    d1, d2, i1, i2 = CO.chamfer_forward(a, b)
    assert np.array_equal(i1, gold[f"{name}_idx1"]) and np.array_equal(i2, gold[f"{name}_idx2"])
    assert np.array_equal(d1, gold[f"{name}_dist1"]) and np.array_equal(d2, gold[f"{name}_dist2"])
    ga, gb = CO.chamfer_backward(a, b, g1, g2, i1, i2)
    np.testing.assert_allclose(ga, gold[f"{name}_grad1"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gb, gold[f"{name}_grad2"], rtol=1e-5, atol=1e-5)


def test_first_minimum_wins():
    a, b, _, _ = CC.case("ties")
    _, _, i1, _ = CO.chamfer_forward(a, b)
    m = b.shape[1]
    assert (i1 < m - m // 2).all()      # never the duplicate in the second half
    d1, _, _, _ = CO.chamfer_forward(a, b)
    k = np.arange(100, 150)                 # a[:50] = t[100:150]; t[128:] duplicates t[:128] -> the earlier copy
    assert (d1[:, :50] == 0).all() and np.array_equal(i1[0, :50], np.where(k < m - m // 2, k, k - (m - m // 2)))
