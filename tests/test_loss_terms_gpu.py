"""GPU: the remaining loss terms of train_step (csrc/loss.cu, csrc/train.cu flow backward) against the CPU
oracle's line-by-line restatement of trainer.py:237-462 — values and gradients (autograd of the oracle)."""
import numpy as np
import pytest

import field_cases as FC

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
from oracle import loss_oracle as LO

S = FC.S


# ---- URF line-of-sight loss --------------------------------------------------------------------------
@pytest.mark.parametrize("R,T,eps", [(1, 8, 0.02), (257, 96, 0.02), (4096, 768, 0.004), (512, 128, 0.002)])
def test_los_loss_matches_oracle(pkg, R, T, eps):
    g = torch.Generator().manual_seed(R + T)
    near, far = S.MIN_NEAR_LIDAR, S.LIDAR_MAX_DEPTH
    z = near + (far - near) * torch.sort(torch.rand(R, T, generator=g), dim=1).values
    w = torch.rand(R, T, generator=g) * 0.2
    gt = near + (far - near) * torch.rand(1, R, generator=g)
    gt[:, ::5] = 0.0                                  # dropped rays (gt_depth * raydrop mask)
    gt[:, 0] = float(z[0, T // 2])                    # a sample exactly at the surface
    w0 = w.clone().requires_grad_(True)
    ref = LO.los_loss(w0, z, gt, eps)
    (3.0 * ref).backward()
    w1 = w.cuda().requires_grad_(True)
    out = pkg.losses.los_loss(w1, z.cuda(), gt.cuda(), eps)
    (3.0 * out).backward()
    assert out.shape == ref.shape == ()
    np.testing.assert_allclose(out.item(), ref.item(), rtol=2e-5)
    np.testing.assert_allclose(w1.grad.cpu().numpy(), w0.grad.numpy(), rtol=2e-5, atol=1e-9)


# ---- structural regularisation -----------------------------------------------------------------------
def _patch_inputs(P, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    H, W = S.LIDAR_H, S.LIDAR_W
    pano = torch.rand(H, W, generator=g) * 0.8
    pano[torch.rand(H, W, generator=g) < 0.1] = 0.0
    # patches of h x w neighbouring pixels, as the dataset samples them
    r0 = torch.randint(0, H - h + 1, (P,), generator=g)
    c0 = torch.randint(0, W - w + 1, (P,), generator=g)
    rows = r0[:, None, None] + torch.arange(h)[None, :, None]
    cols = c0[:, None, None] + torch.arange(w)[None, None, :]
    inds = (rows * W + cols).reshape(1, -1)
    gt_depth = pano.reshape(-1)[inds]
    gt_raydrop = (gt_depth > 0).float()
    pred = (gt_depth + 0.05 * torch.randn(1, P * h * w, generator=g)) * gt_raydrop
    return pano, inds, pred, gt_depth * gt_raydrop, gt_raydrop


@pytest.mark.parametrize("h,w", [(2, 8), (8, 8), (3, 5)])
def test_patch_grad_masks_match_oracle(pkg, h, w):
    pano, inds, _, _, _ = _patch_inputs(64, h, w, 5)
    ex, ey = LO.patch_grad_masks(pano, inds, S.LIDAR_W, h, w, S.SCALE)
    mx, my = pkg.losses.patch_grad_masks(pano.cuda(), inds.cuda(), h, w, S.SCALE)
    assert mx.shape == ex.shape
    assert torch.equal(mx.cpu(), ex.float()) and torch.equal(my.cpu(), ey.float())
    assert 0 < float(ex.float().mean()) < 1     # both classes occur


STRUCT_CASES = [
    dict(grad_loss=True),                                            # the reference's config (kitti360_1908.txt)
    dict(grad_loss=True, sobel_grad=True, depth_grad_loss="mse"),
    dict(grad_loss=True, depth_grad_loss="cos"),
    dict(grad_loss=True, sobel_grad=True, depth_grad_loss="cos", alpha_grad=0.7),
    dict(grad_loss=True, depth_grad_loss="huber"),
    dict(grad_loss=True, depth_grad_loss="smoothl1", sobel_grad=True),
    dict(grad_norm_smooth=True, spatial_smooth=True, tv_loss=True),
    dict(grad_norm_smooth=True, spatial_smooth=True, tv_loss=True, sobel_grad=True, grad_loss=True,
         alpha_grad_norm=0.3, alpha_spatial=0.2, alpha_tv=0.05),
]


@pytest.mark.parametrize("kw", STRUCT_CASES)
@pytest.mark.parametrize("h,w", [(2, 8), (8, 16)])
def test_structural_loss_matches_oracle(pkg, kw, h, w):
    P = 4096 // (h * w)
    pano, inds, pred, gt_depth, gt_raydrop = _patch_inputs(P, h, w, 7)
    ex, ey = LO.patch_grad_masks(pano, inds, S.LIDAR_W, h, w, S.SCALE)
    p0 = pred.clone().requires_grad_(True)
    ref = LO.structural_loss(p0, h, w, S.SCALE, gt_depth, gt_raydrop.unsqueeze(-1), ex, ey, **kw)
    ref.sum().backward()
    p1 = pred.cuda().requires_grad_(True)
    out = pkg.losses.structural_loss(p1, h, w, S.SCALE, gt_depth.cuda(), gt_raydrop.cuda(), ex.float().cuda(),
                                     ey.float().cuda(), **kw)
    out.sum().backward()
    assert out.shape == ref.shape
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=2e-5, atol=1e-6)
    # atol: autograd's cosine backward leaves cancellation residue of ~1e-5 of the largest entry where the
    # closed form is exactly zero
    scale = float(p0.grad.abs().max())
    np.testing.assert_allclose(p1.grad.cpu().numpy(), p0.grad.numpy(), rtol=1e-4, atol=1e-4 * scale)


# ---- scene-flow loss -----------------------------------------------------------------------------------
def _cham(a, b):
    d = ((a[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1)
    m1, m2 = d.min(2), d.min(1)
    return m1.values, m2.values, m1.indices, m2.indices


def _model(pkg):
    m = pkg.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND, min_near=S.MIN_NEAR,
                        min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)
    m.load_flat_params(FC.oracle_params())
    return m.train()


@pytest.mark.parametrize("t,fwd,bwd", [(0.4, True, True), (0.0, True, False), (1.0, False, True)])
def test_flow_loss_matches_oracle(pkg, t, fwd, bwd):
    from oracle.field_oracle import FieldOracle
    rng = np.random.default_rng(3)
    M = 1500
    pc = ((rng.random((M, 3), dtype=np.float32) * 2 - 1) * np.float32(0.8)).astype(np.float32)
    pcf = (pc + 0.01 * rng.standard_normal((M + 17, 3)).astype(np.float32)[:M]).astype(np.float32) if fwd else None
    pcb = (pc[::2] - 0.01).astype(np.float32) if bwd else None
    base = FC.oracle_params()
    leaf = {k: base[k].clone().requires_grad_(True) for k in ("flow_grid", "flow_mlp")}
    params = dict(base, **leaf)
    orc = FieldOracle(FC.oracle_config(), params)
    tt = lambda a: None if a is None else torch.from_numpy(a)
    # The oracle's table gradients pass through an fp16 cast (as tcnn's do); like the reference's GradScaler
    # (trainer.py:119,1332) the loss is scaled by a power of two so that they stay out of the fp16 subnormals.
    ref = LO.flow_loss(lambda x: orc.flow(x, t), _cham, tt(pc), tt(pcf), tt(pcb))
    (ref * FC.LOSS_SCALE).backward()
    # conditioning floor: the oracle's own gradient when the cloud moves by one fp32 ulp
    leaf2 = {k: base[k].clone().requires_grad_(True) for k in ("flow_grid", "flow_mlp")}
    orc2 = FieldOracle(FC.oracle_config(), dict(base, **leaf2))
    (LO.flow_loss(lambda x: orc2.flow(x, t), _cham, tt(np.nextafter(pc, np.float32(10)).astype(np.float32)), tt(pcf),
                  tt(pcb)) * FC.LOSS_SCALE).backward()
    floor = {k: float(torch.linalg.norm(leaf2[k].grad - leaf[k].grad) / torch.linalg.norm(leaf[k].grad)) for k in leaf}
    m = _model(pkg)
    cu = lambda a: None if a is None else torch.from_numpy(a).cuda()
    out = pkg.losses.flow_loss(m, cu(pc), torch.tensor([[t]], device="cuda"), cu(pcf), cu(pcb))
    out.backward()
    assert abs(out.item() - ref.item()) < 1e-2 * abs(ref.item())
    for name in ("flow_grid", "flow_mlp"):
        got = getattr(m, name).grad.detach().cpu().numpy().reshape(-1).astype(np.float64)
        want = leaf[name].grad.numpy().reshape(-1).astype(np.float64) / FC.LOSS_SCALE
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        # flow_grid: every table entry is touched by about one point, so a hidden unit of the flow MLP whose ReLU
        # sign differs between two evaluations (its fp16 input features differ by an fp16 ulp: the kernels
        # interpolate the time-collapsed fp16 table, the oracle interpolates per basis function) changes that
        # entry by ~10 % and nothing averages it out — measured 2.3-2.7 % here, tools/flow_grad_debug.py shows it
        # uniform over the 16 levels and independent of the fp16 gradient scale; flow_mlp sums over all points: 2e-4
        tol = {"flow_grid": 5e-2, "flow_mlp": 1e-2}[name]
        assert err < max(tol, 2.0 * floor[name]), (name, err, floor)
        if name == "flow_grid":
            # ... and the error IS that tail, not a systematic offset: entry by entry the median relative error sits at
            # the oracle's own fp16 resolution (measured 2.8e-4; the oracle's fp16 gradient cast alone gives 3.6e-4),
            # 98 % of the touched entries agree within 1e-2 and the whole rel-L2 error lives in the ~1 % of entries
            # behind a flipped ReLU (tools/flow_grad_debug.py prints the distribution)
            nz = np.abs(want) > 1e-3 * np.abs(want).max()
            r = np.abs(got[nz] - want[nz]) / np.abs(want[nz])
            assert np.median(r) < 2e-3, np.median(r)
            assert float((r > 1e-2).mean()) < 0.05, float((r > 1e-2).mean())
    # nothing but the flow network receives a gradient
    assert m.sigma_net.grad is None and m.hash_static_lidar.grad is None


def test_flow_inference_and_autograd_paths_agree(pkg):
    m = _model(pkg)
    rng = np.random.default_rng(5)
    pc = torch.from_numpy(((rng.random((777, 3), dtype=np.float32) * 2 - 1) * np.float32(1.5))).cuda()
    t = torch.tensor([[0.3]], device="cuda")
    a = m.flow(pc, t)
    with torch.no_grad():
        b = m.flow(pc, t)
    assert a["flow_forward"].requires_grad and not b["flow_forward"].requires_grad
    torch.testing.assert_close(a["flow_forward"].detach(), b["flow_forward"], rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(a["flow_backward"].detach(), b["flow_backward"], rtol=1e-3, atol=1e-6)


def test_forward_only_entry_points_refuse_autograd(pkg):
    m = _model(pkg)
    x = torch.zeros(8, 3, device="cuda")
    with pytest.raises(pkg._lib.NvsfError):
        m.density(x, 0.5, True)
    with torch.no_grad():
        assert m.density(x, 0.5, True)["sigma"].shape == (8,)
