"""Seeded input cases shared by the CPU (oracle vs golden) and GPU (CUDA vs oracle / reference
extension) parity tests and by oracle/make_golden_raymarching.py."""
import importlib

import numpy as np

S = importlib.import_module("selfsupervised-nvsf_b200.synth")

THRESH = 0.01  # renderer_dynamic.py:75 density_thresh default


def bitfield(fill, seed=0):
    return S.packbits_np(S.density_grid(fill, seed=seed), THRESH)


def march_inputs(kind, n, seed, perturb):
    """rays + near/far + noises for the marchers.  kind: 'lidar' | 'camera'."""
    rng = np.random.default_rng(1000 + seed)
    if kind == "lidar":
        o, d = S.lidar_rays(n, seed=seed)
        N = o.shape[0]
        nears = np.full(N, S.MIN_NEAR_LIDAR, np.float32)
        fars = np.full(N, S.LIDAR_MAX_DEPTH, np.float32)
    else:
        o, d = S.camera_rays(n, seed=seed)
        from oracle import raymarching_oracle as O
        nears, fars = O.near_far_from_aabb(o, d, S.AABB, S.MIN_NEAR)
        N = o.shape[0]
    noises = rng.random(N, dtype=np.float32) if perturb else np.zeros(N, np.float32)
    return o, d, nears, fars, noises


def field_values(M, seed):
    """Random per-sample sigmas / rgbs (what the field would return)."""
    rng = np.random.default_rng(2000 + seed)
    sigmas = np.exp(rng.normal(0.0, 2.0, size=M)).astype(np.float32)
    rgbs = rng.random((M, 3), dtype=np.float32)
    return sigmas, rgbs


def canonical_from_rays(rays, arrays):
    """Re-order the per-sample `arrays` of a marcher result into ray-id order.

    rays [N,3] rows (id, offset, count) in ANY order (the reference's order is scheduling
    dependent, raymarching.cu:445-454).  Returns (rays sorted by id with prefix-sum offsets,
    list of arrays concatenated in that order)."""
    rays = np.asarray(rays)
    order = np.argsort(rays[:, 0], kind="stable")
    r = rays[order]
    new_off = np.concatenate([[0], np.cumsum(r[:, 2])[:-1]]).astype(np.int32)
    total = int(r[:, 2].sum())
    idx = np.empty(total, np.int64)
    pos = 0
    for (rid, off, cnt) in r:
        idx[pos:pos + cnt] = np.arange(off, off + cnt)
        pos += cnt
    out_rays = np.stack([r[:, 0], new_off, r[:, 2]], 1).astype(np.int32)
    return out_rays, [np.asarray(a)[idx] for a in arrays]
