"""TEST INFRASTRUCTURE — CPU restatement (numpy, fp32) of the callers either side of the marcher.

* get_lidar_rays / get_rays follow the reference nvsf/nerf/dataset/dataset_utils.py:369-536 and
  :539-687 (direction formulas :512-530 and :667-675).  PINNED: tests/golden/rays_ref.npz holds
  outputs of the reference's own functions imported by file path (oracle/make_golden_rays.py).
* grid_cell_points / grid_update / run_cuda restate torch-ngp's update_extra_state and run_cuda
  around the reference's operators (raymarching.py:85-164, 171-510).  The reference ships no such
  caller and no test for one: PARITY UNPINNED for these three (they are compositions of pinned
  pieces: the raymarching oracle and the field oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

from . import raymarching_oracle as RO

f32 = np.float32


def _pixels(H, W, inds):
    p = np.arange(H * W, dtype=np.int64) if inds is None else np.asarray(inds, np.int64)
    return (p % W).astype(f32), (p // W).astype(f32)


def _emit(dirs, pose):
    pose = np.asarray(pose, f32)
    R = pose[:3, :3]
    d = np.zeros_like(dirs)
    for a in range(3):  # directions @ R^T, left-to-right fp32 sums
        d[:, a] = (dirs[:, 0] * R[a, 0] + dirs[:, 1] * R[a, 1]) + dirs[:, 2] * R[a, 2]
    o = np.broadcast_to(pose[:3, 3], d.shape).copy()
    return o, d


def get_lidar_rays(pose, intrinsics, intrinsics_hoz, H, W, inds=None):
    """dataset_utils.py:512-536 for one pose [4,4]."""
    i, j = _pixels(H, W, inds)
    fov_up, fov = f32(intrinsics[0]), f32(intrinsics[1])
    fov_hoz = f32(intrinsics_hoz[1])
    pi = f32(np.pi)
    beta = -(i - f32(W / 2)) / f32(W) * fov_hoz / f32(180) * pi
    alpha = (fov_up - j / f32(H) * fov) / f32(180) * pi
    dirs = np.stack([np.cos(alpha) * np.cos(beta), np.cos(alpha) * np.sin(beta), np.sin(alpha)], -1).astype(f32)
    return _emit(dirs, pose)


def get_rays(pose, intrinsics, H, W, inds=None):
    """dataset_utils.py:563-681 for one pose [4,4]."""
    i, j = _pixels(H, W, inds)
    K = np.asarray(intrinsics, f32)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    xs = (i + f32(0.5) - cx) / fx
    ys = (j + f32(0.5) - cy) / fy
    nrm = np.sqrt((xs * xs + ys * ys) + f32(1))
    dirs = np.stack([xs / nrm, ys / nrm, f32(1) / nrm], -1).astype(f32)
    return _emit(dirs, pose)


def grid_cell_points(C, H, bound, noise=None):
    """Sample point of every cell in [C][Morton] order (torch-ngp update_extra_state, full update)."""
    H3 = H ** 3
    coords = RO.morton3D_invert(np.arange(H3, dtype=np.int32)).astype(f32)  # [H3,3]
    u = f32(2) * coords / f32(H - 1) - f32(1)
    out = np.empty((C, H3, 3), f32)
    for c in range(C):
        bc = f32(min(2 ** c, bound))
        half = bc / f32(H)
        v = u * (bc - half)
        if noise is not None:
            v = v + (np.asarray(noise, f32).reshape(C, H3, 3)[c] * f32(2) - f32(1)) * half
        out[c] = v
    return out.reshape(C * H3, 3)


def grid_update(density_grid, tmp_grid, decay, density_thresh):
    """-> (new grid, mean, thresh, bitfield)."""
    g = np.asarray(density_grid, f32).reshape(-1).copy()
    t = np.asarray(tmp_grid, f32).reshape(-1)
    valid = (g >= 0) & (t >= 0)
    g[valid] = np.maximum(g[valid] * f32(decay), t[valid])
    mean = f32(np.maximum(g, 0).astype(np.float64).sum() / g.size)
    thresh = f32(min(mean, f32(density_thresh)))
    return g, mean, thresh, RO.packbits(g, thresh)


def run_cuda(field, rays_o, rays_d, t, lidar, bitfield, C, H, bound, nears, fars, dt_gamma=0.0, max_steps=1024,
             T_thresh=1e-4, one_shot=True, noises=None, bg_color=1.0, n_step_fn=None):
    """The march_rays* render loops of torch-ngp's run_cuda over the oracle operators and the
    oracle field (`field` = oracle.field_oracle.FieldOracle)."""
    import torch

    N = rays_o.shape[0]
    nz = np.zeros(N, f32) if noises is None else np.asarray(noises, f32)
    nch = 2 if lidar else 3

    def shade(xyzs, dirs):
        with torch.no_grad():
            r = field.density(torch.from_numpy(xyzs), t, lidar)
            rgb = field.color(torch.from_numpy(dirs), r["geo_feat"], lidar).numpy().astype(f32)
        sig = (r["sigma"].numpy() * f32(field.cfg.density_scale)).astype(f32)
        rgb3 = np.zeros((xyzs.shape[0], 3), f32)
        rgb3[:, :nch] = rgb
        return sig, rgb3

    if one_shot:
        xyzs, dirs, deltas, rays, cnt = RO.march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, nz,
                                                            dt_gamma=dt_gamma, max_steps=max_steps)
        m = int(cnt[0])
        sig, rgb = shade(xyzs[:m], dirs[:m])
        ws, depth, image = RO.composite_rays_train_forward(sig, rgb, deltas[:m], rays, T_thresh)
        # train compositing measures t from the first marching position (raymarching.cu:372-375,626-627)
        dt_min = f32(2 * np.sqrt(3) / max_steps)
        dt_max = f32(2 * np.sqrt(3) * 2 ** (C - 1) / H)
        t0 = np.asarray(nears, f32) + np.clip(np.asarray(nears, f32) * f32(dt_gamma), dt_min, dt_max) * nz
        depth = depth + ws * t0
        n_samples = m
    else:
        ws, depth, image = np.zeros(N, f32), np.zeros(N, f32), np.zeros((N, 3), f32)
        alive = np.arange(N, dtype=np.int32)
        rays_t = np.asarray(nears, f32).copy()
        step, n_samples = 0, 0
        while step < max_steps and alive.size > 0:
            n_alive = alive.size
            n_step = n_step_fn(N, n_alive) if n_step_fn else max(min(N // n_alive, 8), 1)
            noi = nz if step == 0 else np.zeros(n_alive, f32)
            xyzs, dirs, deltas = RO.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, bound, bitfield, C, H,
                                               nears, fars, noi, dt_gamma=dt_gamma, max_steps=max_steps)
            sig, rgb = shade(xyzs, dirs)
            alive, rays_t, ws, depth, image = RO.composite_rays(n_alive, n_step, alive, rays_t, sig, rgb, deltas, ws,
                                                                depth, image, T_thresh)
            alive = alive[alive >= 0]
            n_samples += n_alive * n_step
            step += n_step
    image = image[:, :nch]
    if not lidar:
        image = image + (f32(1) - ws)[:, None] * f32(bg_color)
    return dict(depth=depth, image=image.astype(f32), weights_sum=ws, n_samples=n_samples)
