"""GPU: chamfer_3DDist (csrc/chamfer.cu through the C ABI) against the C oracle, against the golden
vectors of the reference's own extension, against that extension live when its .so travelled
(oracle/_ref/chamfer_3D_ref.so), and — at the 64 K-point size of BASELINE configs[4] — through
properties: a permuted copy of a cloud has distance 0 and the inverse permutation as index."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import chamfer_cases as CC
from conftest import GOLDEN, ROOT
from oracle import chamfer_oracle as CO

REF_SO = os.path.join(ROOT, "oracle", "_ref", "chamfer_3D_ref.so")


def run(pkg, a, b, g1, g2):
    xa = torch.from_numpy(a).cuda().requires_grad_(True)
    xb = torch.from_numpy(b).cuda().requires_grad_(True)
    d1, d2, i1, i2 = pkg.chamfer.chamfer_3DDist()(xa, xb)
    ((d1 * torch.from_numpy(g1).cuda()).sum() + (d2 * torch.from_numpy(g2).cuda()).sum()).backward()
    return [t.detach().cpu().numpy() for t in (d1, d2, i1, i2, xa.grad, xb.grad)]


@pytest.mark.parametrize("name", list(CC.GOLDEN_CASES))
def test_matches_oracle_and_reference_golden(pkg, name):
    a, b, g1, g2 = CC.case(name)
    d1, d2, i1, i2, ga, gb = run(pkg, a, b, g1, g2)
    od1, od2, oi1, oi2 = CO.chamfer_forward(a, b)
    assert i1.dtype == np.int32 and np.array_equal(i1, oi1) and np.array_equal(i2, oi2)
    assert np.array_equal(d1, od1) and np.array_equal(d2, od2)
    oga, ogb = CO.chamfer_backward(a, b, g1, g2, oi1, oi2)
    np.testing.assert_allclose(ga, oga, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gb, ogb, rtol=1e-5, atol=1e-5)
    path = os.path.join(GOLDEN, "chamfer_ref_sm100a.npz")
    if os.path.exists(path):
        gold = np.load(path)
        assert np.array_equal(i1, gold[f"{name}_idx1"]) and np.array_equal(i2, gold[f"{name}_idx2"])
        assert np.array_equal(d1, gold[f"{name}_dist1"]) and np.array_equal(d2, gold[f"{name}_dist2"])
        np.testing.assert_allclose(ga, gold[f"{name}_grad1"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(gb, gold[f"{name}_grad2"], rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="reference extension not built (oracle/build_ref_chamfer.sh)")
@pytest.mark.parametrize("n,m", [(8192, 8192), (20000, 3001)])
def test_matches_reference_extension_live(pkg, n, m):
    spec = importlib.util.spec_from_file_location("chamfer_3D_ref", REF_SO)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(n)
    a = (rng.random((1, n, 3), dtype=np.float32) * 2 - 1).astype(np.float32)
    b = (rng.random((1, m, 3), dtype=np.float32) * 2 - 1).astype(np.float32)
    xa, xb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    d1, d2 = torch.zeros(1, n).cuda(), torch.zeros(1, m).cuda()
    i1, i2 = torch.zeros(1, n, dtype=torch.int32).cuda(), torch.zeros(1, m, dtype=torch.int32).cuda()
    ref.forward(xa, xb, d1, d2, i1, i2)
    o1, o2, j1, j2 = pkg.chamfer.chamfer_3DDist()(xa, xb)
    assert torch.equal(j1, i1) and torch.equal(j2, i2) and torch.equal(o1, d1) and torch.equal(o2, d2)


def test_permuted_cloud_at_64k_points(pkg):
    n = 65536
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.rand(1, n, 3, device="cuda", generator=g)
    perm = torch.randperm(n, device="cuda", generator=g)
    b = a[:, perm]
    d1, d2, i1, i2 = pkg.chamfer.chamfer_3DDist()(a, b)
    assert float(d1.max()) == 0.0 and float(d2.max()) == 0.0
    assert torch.equal(i2[0].long(), perm)                       # b[k] = a[perm[k]]
    inv = torch.empty_like(perm); inv[perm] = torch.arange(n, device="cuda")
    assert torch.equal(i1[0].long(), inv)


def test_shape_errors_and_cpu_tensors(pkg):
    with pytest.raises(AssertionError):
        pkg.chamfer.chamfer_3DDist()(torch.zeros(1, 4, 2).cuda(), torch.zeros(1, 4, 3).cuda())
    with pytest.raises(pkg._lib.NvsfError):
        pkg.chamfer.chamfer_3DDist()(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
