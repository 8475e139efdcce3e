"""Training plumbing on the GPU: the flat Adam kernel against torch.optim.Adam, and a few joint
LiDAR + camera steps through NeRFNetwork.render -> loss -> backward -> GradSync -> FlatAdam."""
import importlib

import numpy as np
import pytest

import field_cases as FC

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
S = FC.S
SMALL = dict(log2_hashmap_size=12, hash_size_dynamic=(10, 9, 9), flow_log2_hashmap_size=11,
             time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND, min_near=S.MIN_NEAR,
             min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)


def test_flat_adam_matches_torch_adam(pkg):
    torch.manual_seed(0)
    m = pkg.NeRFNetwork(**SMALL).train()
    ref = {n: p.detach().clone().requires_grad_(True) for n, p in m.named_parameters()}
    lr_scale = pkg.optim.LR_SCALE
    topt = torch.optim.Adam([{"params": [p], "lr": 1e-2 * lr_scale.get(n, 1.0)} for n, p in ref.items()],
                            betas=(0.9, 0.99), eps=1e-15)
    opt = pkg.optim.FlatAdam(m, lr=1e-2)
    for p in m.parameters():   # parameters and gradients are views of the flat buffers
        assert p.data.untyped_storage().data_ptr() == opt.flat.untyped_storage().data_ptr()
    for it in range(3):
        g = torch.Generator(device="cuda").manual_seed(it)
        for n, p in m.named_parameters():
            grad = torch.randn(p.shape, generator=g, device="cuda") * (10.0 ** (it - 1))
            grad[::3] = 0.0       # untouched table entries: zero gradient, moments still decay
            p.grad.copy_(grad)
            ref[n].grad = grad.clone()
        opt.step()
        topt.step()
    for n, p in m.named_parameters():
        np.testing.assert_allclose(p.detach().cpu().numpy(), ref[n].detach().cpu().numpy(), rtol=2e-5, atol=1e-7,
                                   err_msg=n)


def test_nonfinite_gradient_skips_the_step_on_the_device(pkg):
    """scaler.step(optimizer) semantics (trainer.py:1332-1334) with no host read: a step whose gradient holds an
    inf / nan leaves parameters, moments and the applied-step count untouched; the next finite step equals
    torch.optim.Adam's FIRST step (bias corrections follow the applied count, not the requested one)."""
    torch.manual_seed(0)
    m = pkg.NeRFNetwork(**SMALL).train()
    ref = {n: p.detach().clone().requires_grad_(True) for n, p in m.named_parameters()}
    topt = torch.optim.Adam([{"params": [p], "lr": 1e-2 * pkg.optim.LR_SCALE.get(n, 1.0)} for n, p in ref.items()],
                            betas=(0.9, 0.99), eps=1e-15)
    opt = pkg.optim.FlatAdam(m, lr=1e-2, skip_nonfinite=True)
    before = opt.flat.clone()
    g = torch.Generator(device="cuda").manual_seed(5)
    for bad in (float("inf"), float("nan")):
        opt.sync.flat.copy_(torch.randn(opt.sync.flat.shape, generator=g, device="cuda"))
        m.sigma_net.grad.view(-1)[-1] = bad      # the very last element of a tensor in the middle of the buffer
        opt.step()
        assert torch.equal(opt.flat, before) and float(opt.exp_avg.abs().sum()) == 0.0
        assert float(opt.found_inf) == 1.0 and opt.applied_steps() == 0
    for n, p in m.named_parameters():
        grad = torch.randn(p.shape, generator=g, device="cuda")
        p.grad.copy_(grad)
        ref[n].grad = grad.clone()
    opt.step()
    topt.step()
    assert opt.applied_steps() == 1 and float(opt.found_inf) == 0.0
    for n, p in m.named_parameters():
        np.testing.assert_allclose(p.detach().cpu().numpy(), ref[n].detach().cpu().numpy(), rtol=2e-5, atol=1e-7,
                                   err_msg=n)
    L = pkg._lib.lib()
    x = torch.zeros(16, device="cuda")
    assert L.nvsf_grad_found_inf(x.data_ptr() + 4, 8, x.data_ptr(), None) == -1      # misaligned
    assert L.nvsf_grad_found_inf(x.data_ptr(), 6, x.data_ptr(), None) == -1          # not a multiple of 4
    assert L.nvsf_adam_begin(None, None, 0.9, 0.99, None) == -1


def test_adam_rejects_bad_arguments(pkg):
    L = pkg._lib.lib()
    x = torch.zeros(16, device="cuda")
    assert L.nvsf_adam_step(x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), 16, 1e-2, 0.9, 0.99, 1e-15, 0, 1.0,
                            None) == -1      # step is 1-based
    assert L.nvsf_adam_step(x.data_ptr() + 4, x.data_ptr(), x.data_ptr(), x.data_ptr(), 8, 1e-2, 0.9, 0.99, 1e-15, 1,
                            1.0, None) == -1  # misaligned


def test_joint_training_steps_reduce_the_loss(pkg):
    """LiDAR depth/intensity + camera colour regression on fixed synthetic targets: the loss of
    a handful of Adam steps must go down, every gradient stays finite, and the re-packed tables
    follow the updated parameters."""
    torch.manual_seed(0)
    m = pkg.NeRFNetwork(**SMALL).train()
    opt = pkg.optim.FlatAdam(m, lr=5e-3)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    lo, ld = map(dev, S.lidar_rays(256, seed=1))
    co, cd = map(dev, S.camera_rays(256, seed=2))
    t = torch.tensor([[0.4]], device="cuda")
    gt_d = torch.full((256,), 0.3, device="cuda")
    gt_c = torch.rand(256, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    losses = []
    for it in range(12):
        opt.zero_grad()
        ol = m.render(lo[None], ld[None], t, cal_lidar_color=True, staged=False, num_steps=64, perturb=True)
        l1 = (ol["depth_lidar"].view(-1) - gt_d).abs().mean() + (ol["image_lidar"] ** 2).mean()
        l1.backward()
        opt.sync.reduce_group("lidar")
        oc = m.render(co[None], cd[None], t, cal_lidar_color=False, staged=False, num_steps=64, perturb=True)
        l2 = ((oc["image"].view(-1, 3) - gt_c) ** 2).mean()
        l2.backward()
        opt.sync.reduce_group("camera"); opt.sync.reduce_group("shared"); opt.sync.wait()
        assert torch.isfinite(opt.sync.flat).all()
        opt.step()
        losses.append(float(l1) + float(l2))
    assert losses[-1] < 0.7 * losses[0], losses
