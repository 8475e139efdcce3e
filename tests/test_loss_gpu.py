"""GPU: the loss head (csrc/loss.cu through selfsupervised-nvsf_b200/losses.py) against the CPU
oracle — values to 1e-6, gradients (autograd of the oracle) to 1e-6, all criteria, empty input."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
from oracle import loss_oracle as LO

CASES = [dict(), dict(smooth_factor=0.2), dict(depth_loss="mse", raydrop_loss="l1", intensity_loss="l1"),
         dict(depth_loss="smoothl1", raydrop_loss="smoothl1", intensity_loss="huber", scale=0.7, alpha_d=0.5,
              alpha_r=0.3, alpha_i=2.0),
         dict(depth_loss="huber", raydrop_loss="huber", intensity_loss="smoothl1", scale=0.010851959895748291)]


def _inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    depth = torch.rand(1, n, generator=g)
    image = torch.rand(1, n, 2, generator=g)
    gt = torch.rand(1, n, 3, generator=g)
    gt[:, :, 0] = (gt[:, :, 0] > 0.3).float()
    return depth, image, gt


@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("n", [1, 4096, 67980])
def test_lidar_loss_matches_oracle(pkg, kw, n):
    depth, image, gt = _inputs(n, 3 + n)
    d0, i0 = depth.clone().requires_grad_(True), image.clone().requires_grad_(True)
    ref = LO.lidar_loss(d0, i0, gt, **kw)
    w = torch.rand(1, n, generator=torch.Generator().manual_seed(9))
    (ref * w).sum().backward()
    d1, i1 = depth.cuda().requires_grad_(True), image.cuda().requires_grad_(True)
    out = pkg.losses.lidar_loss(d1, i1, gt.cuda(), **kw)
    assert out.shape == ref.shape
    (out * w.cuda()).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(d1.grad.cpu().numpy(), d0.grad.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(i1.grad.cpu().numpy(), i0.grad.numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["mse", "l1", "smoothl1", "huber"])
def test_rgb_loss_matches_oracle(pkg, name):
    g = torch.Generator().manual_seed(11)
    p, t = torch.rand(1, 4096, 3, generator=g), torch.rand(1, 4096, 3, generator=g)
    p0 = p.clone().requires_grad_(True)
    ref = LO.rgb_loss(p0, t, alpha_rgb=1.5, rgb_loss=name, scale=0.5)
    ref.sum().backward()
    p1 = p.cuda().requires_grad_(True)
    out = pkg.losses.rgb_loss(p1, t.cuda(), alpha_rgb=1.5, rgb_loss=name, scale=0.5)
    out.sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(p1.grad.cpu().numpy(), p0.grad.numpy(), rtol=1e-6, atol=1e-7)


def test_unbuilt_criterion_and_empty_input(pkg):
    with pytest.raises(pkg._lib.NvsfError):
        pkg.losses.rgb_loss(torch.zeros(1, 4, 3).cuda(), torch.zeros(1, 4, 3).cuda(), rgb_loss="bce")
    e = pkg.losses.lidar_loss(torch.zeros(1, 0).cuda(), torch.zeros(1, 0, 2).cuda(), torch.zeros(1, 0, 3).cuda())
    assert e.shape == (1, 0)
