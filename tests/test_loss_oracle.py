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
