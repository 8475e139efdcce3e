"""CPU: the loss oracle (oracle/loss_oracle.py) against hand-evaluated cases of the reference's
train_step supervision (trainer.py:188-219, 503-504) and its autograd against closed forms."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from oracle import loss_oracle as LO


def test_lidar_loss_hand_case():
    # one dropped ray (mask 0), one kept ray; defaults alpha 1 / 0.01 / 0.1, l1 / mse / mse
    depth = torch.tensor([[0.5, 0.25]])
    image = torch.tensor([[[0.75, 0.5], [0.25, 1.0]]])       # (raydrop, intensity)
    gt = torch.tensor([[[1.0, 0.25, 0.75], [0.0, 0.5, 0.5]]])  # (mask, intensity, depth)
    got = LO.lidar_loss(depth, image, gt).numpy()
    kept = 1.0 * abs(0.5 - 0.75) + 0.01 * (0.75 - 1.0) ** 2 + 0.1 * (0.5 - 0.25) ** 2
    dropped = 0.0 + 0.01 * (0.25 - 0.0) ** 2 + 0.0
    np.testing.assert_allclose(got, [[kept, dropped]], rtol=1e-6)


def test_label_smoothing_and_criteria():
    depth = torch.tensor([[0.4]])
    image = torch.tensor([[[0.9, 0.3]]])
    gt = torch.tensor([[[1.0, 0.1, 0.1]]])
    got = LO.lidar_loss(depth, image, gt, smooth_factor=0.2, depth_loss="smoothl1", raydrop_loss="l1",
                        intensity_loss="huber", scale=0.5).item()
    d = 0.3                                     # |0.4 - 0.1| >= beta 0.1 -> d - 0.05
    want = (d - 0.05) + 0.01 * abs(0.9 - 0.8) + 0.1 * (0.1 * (0.2 - 0.05))  # huber delta 0.1, |0.2| > delta
    assert abs(got - want) < 1e-6


def test_rgb_loss_and_gradient():
    p = torch.tensor([[[0.2, 0.5, 0.9]]], requires_grad=True)
    g = torch.tensor([[[0.1, 0.5, 1.0]]])
    l = LO.rgb_loss(p, g, alpha_rgb=2.0)
    np.testing.assert_allclose(l.detach().numpy(), 2.0 * (p.detach().numpy() - g.numpy()) ** 2, rtol=1e-6)
    l.sum().backward()
    np.testing.assert_allclose(p.grad.numpy(), 4.0 * (p.detach().numpy() - g.numpy()), rtol=1e-6)


# ---- hand-evaluated cases for the remaining terms (trainer.py:237-462) ---------------------------------
def test_los_loss_by_hand():
    # one ray, gt depth 0.5, eps 0.1: samples at 0.2 (empty), 0.45 (near), 0.5 (at the surface), 0.9 (empty)
    z = torch.tensor([[0.2, 0.45, 0.5, 0.9]])
    w = torch.tensor([[0.1, 0.3, 0.5, 0.05]])
    gt = torch.tensor([[0.5]])
    eps, sigma = 0.1, 0.1 / 3
    distr = [0.0, np.exp(-(0.05 ** 2) / (2 * sigma ** 2)), 1.0, 0.0]   # normalised by its maximum (the peak)
    empty = 0.1 ** 2 + 0.05 ** 2
    near = (0.0 - distr[0]) ** 2 + (0.3 - distr[1]) ** 2 + (0.5 - 1.0) ** 2 + 0.0
    want = 0.1 * empty / 1 + 0.1 * near / 1
    np.testing.assert_allclose(LO.los_loss(w, z, gt, eps).item(), want, rtol=1e-5)


def test_structural_loss_by_hand():
    # one 2 x 2 patch, scale 1: d = [[1, 3], [2, 7]]; finite differences with the last column / row repeated
    pred = torch.tensor([[1.0, 3.0, 2.0, 7.0]])
    gx = np.array([[-2.0, -2.0], [-5.0, -5.0]])
    gy = np.array([[-1.0, -4.0], [-1.0, -4.0]])
    out = LO.structural_loss(pred, 2, 2, 1.0, tv_loss=True, spatial_smooth=True, alpha_tv=0.5,
                                      alpha_spatial=2.0)
    want = 0.5 * (np.abs(gx) + np.abs(gy)) + 2.0 * (gx ** 2 + gy ** 2)
    np.testing.assert_allclose(out.numpy().reshape(2, 2), want, rtol=1e-6)
    # gradient loss, l1, all masks on, gt = 0: alpha * sum(|gx| + |gy|)
    ones = torch.ones(1, 1, 2, 2)
    gl = LO.structural_loss(pred, 2, 2, 1.0, gt_depth=torch.zeros(1, 4), gt_raydrop=torch.ones(1, 4, 1),
                                     grad_mask_x=ones, grad_mask_y=ones, grad_loss=True, alpha_grad=0.1)
    np.testing.assert_allclose(gl.item(), 0.1 * (np.abs(gx).sum() + np.abs(gy).sum()), rtol=1e-6)
    # Sobel at the centre of a 3 x 3 ramp d(y, x) = x: gx = 8, gy = 0
    ramp = torch.tensor([[0.0, 1.0, 2.0] * 3])
    sx, sy = LO._grads(ramp.reshape(1, 1, 3, 3), True)
    assert sx[0, 0, 1, 1].item() == 8.0 and sy[0, 0, 1, 1].item() == 0.0
