"""GPU parity at BASELINE.json's FULL configurations, with "trained"-style tables (oracle/field_init.py:
tables U(-1,1), perturbed time planes, a flow that moves the warped queries) so that sigma is not ~1
everywhere:

  configs[1]  66 x 1030 LiDAR frame x 768 samples, render(staged=True)      vs the field oracle on a strided ray sample
  configs[3]  376 x 1408 camera frame with occupancy skipping (run_cuda)    sample offsets / counts / positions of ALL
              529 408 rays bit-exact vs the C oracle; image vs the scene oracle on a ray sample
  configs[2]  4096 + 4096 rays x 768 samples training step                  per-tensor gradient rel-L2 vs the oracle's
              autograd on a 256-ray sub-batch (the loss reads only those rays, the launch is the full batch)

Tolerances: 1e-2 relative for outputs of the fp16 MLPs (BASELINE.json north_star); bit-exact sampling."""
import numpy as np
import pytest

import field_cases as FC
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
S = FC.S


def host(t):
    return t.detach().float().cpu().numpy()


def close(a, b, rtol, atol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), f"{what}: worst excess {err.max():.3e}, max abs err {np.abs(a - b).max():.3e}"


def make_model(pkg, train=False):
    m = pkg.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                        min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)
    m.load_flat_params(FC.oracle_params())
    return m.train() if train else m.eval()


@pytest.fixture(scope="module")
def orc():
    from oracle.field_oracle import FieldOracle
    return FieldOracle(FC.oracle_config(), FC.oracle_params())


def test_lidar_frame_768_samples_vs_oracle(pkg, orc):
    """BASELINE configs[1] at full size through the public staged render."""
    model = make_model(pkg)
    o, d = S.lidar_rays(-1, seed=0)
    assert o.shape[0] == 66 * 1030
    to, td = torch.from_numpy(o).cuda()[None], torch.from_numpy(d).cuda()[None]
    t = torch.tensor([[31.0 / 63.0]], device="cuda")
    full = model.render(to, td, t, cal_lidar_color=True, staged=True, num_steps=768)
    depth, img = host(full["depth_lidar"])[0], host(full["image_lidar"])[0]
    assert depth.shape == (67980,) and img.shape == (67980, 2)
    assert np.isfinite(depth).all() and np.isfinite(img).all() and (img >= 0).all() and (img <= 1).all()
    idx = np.arange(331, 67980, 709)                     # 96 rays over all rows and azimuths
    with torch.no_grad():
        e = orc.run(torch.from_numpy(o[idx]), torch.from_numpy(d[idx]), 31.0 / 63.0, True, 768)
    assert 0.05 < float(e["weights_sum"].mean()) < 0.95 and float(e["sigma"].std()) > 0.01   # not the flat init field
    close(depth[idx], e["depth"].numpy(), 1e-2, 1e-5, "depth")
    close(img[idx], e["image"].numpy(), 1e-2, 1e-4, "image")
    # size-independent property: any sub-range of rays rendered alone gives the same pixels
    with torch.no_grad():
        sub = model.render(to[:, 40000:44096], td[:, 40000:44096], t, cal_lidar_color=True, staged=False, num_steps=768)
    np.testing.assert_allclose(host(sub["depth_lidar"])[0], depth[40000:44096], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(host(sub["image_lidar"])[0], img[40000:44096], rtol=1e-6, atol=1e-7)


def test_camera_frame_march_bit_exact_and_image_vs_oracle(pkg, orc, oracle):
    """BASELINE configs[3] at full size: all 529 408 rays of the 376 x 1408 frame on the street-shell grid."""
    from oracle import scene_oracle as SO
    model = make_model(pkg)
    rm = pkg.raymarching
    o, d = S.camera_rays(-1, seed=0)
    N = o.shape[0]
    assert N == 376 * 1408
    bf = S.packbits_np(S.density_grid("shell"), 0.01)
    noises = np.random.default_rng(4).random(N, dtype=np.float32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    to, td, tbf, tno = dev(o), dev(d), dev(bf), dev(noises)
    nears, fars = rm.near_far_from_aabb(to, td, dev(S.AABB), S.MIN_NEAR)
    en, ef = oracle.near_far_from_aabb(o, d, S.AABB, S.MIN_NEAR)
    assert_bits_equal(host(nears), en, "nears")
    assert_bits_equal(host(fars), ef, "fars")
    xyzs, dirs, deltas, rays = rm.march_rays_train(to, td, S.BOUND, tbf, S.CASCADE, S.GRID_SIZE, nears, fars, None, -1,
                                                   True, -1, True, S.DT_GAMMA, 1024, tno)
    # the oracle sizes N * max_steps rows by default (like the reference wrapper): count first, then M rows
    m = int(oracle.march_rays_train(o, d, S.BOUND, bf, S.CASCADE, S.GRID_SIZE, en, ef, noises, dt_gamma=S.DT_GAMMA,
                                    M=1)[4][0])
    ex, ed, el, er, ec = oracle.march_rays_train(o, d, S.BOUND, bf, S.CASCADE, S.GRID_SIZE, en, ef, noises,
                                                 dt_gamma=S.DT_GAMMA, M=m)
    assert m > 10_000_000 and xyzs.shape[0] == m and int(ec[0]) == m
    assert np.array_equal(rays.cpu().numpy(), er), "ray ids / sample offsets / counts"
    assert_bits_equal(xyzs.cpu().numpy(), ex[:m], "xyzs")
    assert_bits_equal(deltas.cpu().numpy(), el[:m], "deltas")
    assert_bits_equal(dirs.cpu().numpy(), ed[:m], "dirs")
    del xyzs, dirs, deltas, ex, ed, el
    # the rendered frame (public API) against the scene oracle on a strided ray sample
    t = 0.5
    kw = dict(dt_gamma=S.DT_GAMMA, max_steps=1024, T_thresh=1e-2, one_shot=True)
    r = model.run_cuda(to[None], td[None], t, cal_lidar_color=False, noises=tno, density_bitfield=tbf, perturb=True, **kw)
    assert model.last_run_cuda_samples == m
    img, dep, ws = host(r["image"])[0], host(r["depth"])[0], host(r["weights_sum"])
    assert img.shape == (N, 3) and np.isfinite(img).all()
    idx = np.arange(1000, N, 2753)                       # 192 rays
    e = SO.run_cuda(orc, o[idx], d[idx], t, False, bf, S.CASCADE, S.GRID_SIZE, S.BOUND, en[idx], ef[idx],
                    noises=noises[idx], **kw)
    assert e["weights_sum"].max() > 0.05
    close(ws[idx], e["weights_sum"], 1e-2, 1e-5, "weights_sum")
    close(dep[idx], e["depth"], 1e-2, 1e-5, "depth")
    close(img[idx], e["image"], 1e-2, 1e-4, "image")


# Relative L2 error allowed per gradient tensor.  The reference gradient of this field is itself only defined to
# 1-3 %: tests/test_field_grad_gpu.py measures how far the CPU oracle's own gradient moves when the ray origins
# change by ONE fp32 ulp (FC.oracle_grad_floor: 0.8-2 % for tables and MLPs, up to 2.6 % for the flow tensors — the
# finest hash level has 32768 cells per unit, every MLP stores fp16 activations) and allows twice that floor; the
# fixed numbers below are that rule at this test's size, where a second oracle pass would cost minutes.
GRAD_RTOL = {"hash_static": 3e-2, "hash_dynamic": 3e-2, "planes": 3e-2, "flow_grid": 6e-2, "flow_mlp": 3e-2,
             "sigma_net": 2e-2, "intensity_net": 2e-2, "raydrop_net": 2e-2, "color_net": 2e-2}


@pytest.mark.parametrize("lidar", [True, False])
def test_train_step_gradients_4096_rays_768_samples(pkg, lidar):
    """BASELINE configs[2]: one modality's half of the joint step at full size (4096 rays x 768 samples,
    perturb=True); the loss reads 256 rays spread over the batch, whose oracle gradient is affordable."""
    from oracle import raymarching_oracle as RO
    from oracle.field_oracle import FieldOracle
    N, Sn, t = 4096, 768, 0.4
    o, d = (S.lidar_rays if lidar else S.camera_rays)(N, seed=21)
    rng = np.random.default_rng(8)
    noise = rng.random((N, Sn), dtype=np.float32)
    sub = np.arange(7, N, 16)                             # 256 rays
    nch = 2 if lidar else 3
    ca = rng.normal(size=sub.size).astype(np.float32)
    cb = rng.normal(size=(sub.size, nch)).astype(np.float32)
    ce = (rng.normal(size=sub.size) * 0.5).astype(np.float32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    m = make_model(pkg, train=True)
    out = m.render(dev(o)[None], dev(d)[None], torch.tensor([[t]], device="cuda"), cal_lidar_color=lidar, staged=False,
                   num_steps=Sn, noise=dev(noise))
    sfx = "_lidar" if lidar else ""
    ts = torch.from_numpy(sub).cuda()
    loss = ((dev(ca) * out["depth" + sfx].reshape(-1)[ts]).sum() + (dev(cb) * out["image" + sfx].reshape(N, nch)[ts]).sum()
            + (dev(ce) * out["weights_sum" + sfx][ts]).sum())
    loss.backward()
    torch.cuda.synchronize()
    # oracle: the same 256 rays alone (rays are independent)
    base = FC.oracle_params()
    mod = "lidar" if lidar else "camera"
    leaf = {mod: {k: v.clone().requires_grad_(True) for k, v in base[mod].items()}}
    for k in FC.GRAD_SHARED:
        leaf[k] = base[k].clone().requires_grad_(True)
    orc = FieldOracle(FC.oracle_config(), leaf)
    nears = fars = None
    if not lidar:
        n_, f_ = RO.near_far_from_aabb(o[sub], d[sub], S.AABB, S.MIN_NEAR)
        nears, fars = torch.from_numpy(n_), torch.from_numpy(f_)
    e = orc.run(torch.from_numpy(o[sub]), torch.from_numpy(d[sub]), t, lidar, Sn, nears, fars, torch.from_numpy(noise[sub]))
    eloss = ((torch.from_numpy(ca) * e["depth"]).sum() + (torch.from_numpy(cb) * e["image"]).sum()
             + (torch.from_numpy(ce) * e["weights_sum"]).sum())
    (eloss * FC.LOSS_SCALE).backward()   # GradScaler-like: the oracle's table gradients pass through an fp16 cast
    assert abs(loss.item() - eloss.item()) < 1e-2 * max(1.0, abs(eloss.item()))
    close(host(out["depth" + sfx]).reshape(-1)[sub], e["depth"].detach().numpy(), 1e-2, 1e-5, "depth")
    errs = {}
    for name in FC.GRAD_NAMES:
        p = getattr(m, f"{name}_{mod}" if name in ("hash_static", "hash_dynamic", "planes") else name)
        lp = leaf[mod][name] if name in leaf[mod] else leaf[name]
        ref = (lp.grad if lp.grad is not None else torch.zeros_like(lp)).numpy().reshape(-1).astype(np.float64) / FC.LOSS_SCALE
        got = np.zeros_like(ref) if p.grad is None else p.grad.detach().cpu().numpy().reshape(-1).astype(np.float64)
        if not ref.any():
            assert not got.any(), name
            continue
        errs[name] = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    print("gradient rel-L2 errors", {k: round(v, 5) for k, v in errs.items()})
    for name, err in errs.items():
        assert err < GRAD_RTOL[name], (name, err, errs)
