"""Seeded synthetic inputs shaped like the reference's KITTI-360 data (numpy only).

The dataset is not available offline, so tests and bench.py use rays generated
with the same geometry as the reference's ray generators
(nvsf/nerf/dataset/dataset_utils.py:512-530 `get_lidar_rays`, :563-681 `get_rays`)
and the scene scalars of nvsf/configs/kitti360_1908.txt / scripts/main_nvsf.py.
"""
import math

import numpy as np

# scene scalars (kitti360_1908.txt:5-9, main_nvsf.py:28-35,47,116,167-169)
SCALE = 0.010851959895748291
BOUND = 2.0
MIN_NEAR = 1.0 * SCALE
MIN_NEAR_LIDAR = 1.0 * SCALE
LIDAR_MAX_DEPTH = 80.0 * SCALE
NUM_FRAMES = 64
TIME_RESOLUTION = 8
DT_GAMMA = 1.0 / 128
MAX_STEPS = 1024
GRID_SIZE = 128
CASCADE = 1 + math.ceil(math.log2(BOUND))  # renderer_dynamic.py:82
LIDAR_H, LIDAR_W = 66, 1030                # preprocess_data.py:22-31
LIDAR_FOV_UP, LIDAR_FOV = 2.0, 26.9        # intrinsics_lidar
LIDAR_FOV_HOZ = 360.0
CAM_H, CAM_W = 376, 1408
CAM_FX = CAM_FY = 552.554261
CAM_CX, CAM_CY = 682.049453, 238.769549
AABB = np.array([-BOUND, -BOUND, -BOUND, BOUND, BOUND, BOUND], dtype=np.float32)


def random_pose(seed):
    """Random yaw + translation U(-0.3, 0.3)^3 (the scene is normalised to about +-1)."""
    rng = np.random.default_rng(seed)
    yaw = rng.uniform(-math.pi, math.pi)
    c, s = math.cos(yaw), math.sin(yaw)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=np.float64)
    t = rng.uniform(-0.3, 0.3, size=3)
    return R, t


def lidar_rays(n=-1, seed=0, H=LIDAR_H, W=LIDAR_W):
    """LiDAR range-image rays.  n<=0: all H*W pixels row-major; else n random pixel ids."""
    R, t = random_pose(seed)
    rng = np.random.default_rng(seed + 1)
    inds = np.arange(H * W) if n <= 0 else rng.integers(0, H * W, size=n)
    i = (inds % W).astype(np.float64)   # column
    j = (inds // W).astype(np.float64)  # row
    beta = -(i - W / 2) / W * LIDAR_FOV_HOZ / 180.0 * np.pi
    alpha = (LIDAR_FOV_UP - j / H * LIDAR_FOV) / 180.0 * np.pi
    d = np.stack([np.cos(alpha) * np.cos(beta), np.cos(alpha) * np.sin(beta), np.sin(alpha)], -1)
    rays_d = (d @ R.T).astype(np.float32)
    rays_o = np.broadcast_to(t.astype(np.float32), rays_d.shape).copy()
    return rays_o, rays_d


def camera_rays(n=-1, seed=0, H=CAM_H, W=CAM_W):
    """Pinhole camera rays with the KITTI-360 cam_00 intrinsics (x right, y down, z forward)."""
    R, t = random_pose(seed + 100)
    # camera looks along world +x after the yaw: columns = (right, down, forward) axes
    cam2world = R @ np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    rng = np.random.default_rng(seed + 101)
    inds = np.arange(H * W) if n <= 0 else rng.integers(0, H * W, size=n)
    i = (inds % W).astype(np.float64) + 0.5
    j = (inds // W).astype(np.float64) + 0.5
    d = np.stack([(i - CAM_CX) / CAM_FX, (j - CAM_CY) / CAM_FY, np.ones_like(i)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays_d = (d @ cam2world.T).astype(np.float32)
    rays_o = np.broadcast_to(t.astype(np.float32), rays_d.shape).copy()
    return rays_o, rays_d


def _morton3d(x, y, z):
    def spread(v):
        v = v.astype(np.uint32) & 0x3FF
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    return spread(x) | (spread(y) << 1) | (spread(z) << 2)


def density_grid(fill, seed=0, C=CASCADE, H=GRID_SIZE):
    """Synthetic density grid [C, H^3] (Morton order inside each cascade), float32.

    fill: 'full' (all occupied), 'random5' (5 % of cells), 'shell' (cells within two
    voxels of the planes z=-0.05 and |y|=0.3 — a street-like scene), 'empty'.
    Occupied cells hold 1.0, empty cells 0.0.
    """
    grid = np.zeros((C, H * H * H), dtype=np.float32)
    if fill == "full":
        grid[:] = 1.0
    elif fill == "empty":
        pass
    elif fill == "random5":
        rng = np.random.default_rng(seed)
        grid[rng.random(grid.shape) < 0.05] = 1.0
    elif fill == "shell":
        ax = np.arange(H)
        X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
        idx = _morton3d(X.ravel(), Y.ravel(), Z.ravel())
        for c in range(C):
            mip_bound = min(2.0 ** c, BOUND)
            half = mip_bound / H  # half voxel
            cx = ((X.ravel() + 0.5) / H * 2 - 1) * mip_bound
            cy = ((Y.ravel() + 0.5) / H * 2 - 1) * mip_bound
            cz = ((Z.ravel() + 0.5) / H * 2 - 1) * mip_bound
            tol = 4 * half
            occ = (np.abs(cz + 0.05) <= tol) | (np.abs(np.abs(cy) - 0.3) <= tol)
            del cx
            grid[c, idx[occ]] = 1.0
    else:
        raise ValueError(fill)
    return grid


def packbits_np(grid, thresh):
    """numpy statement of packbits: bit i of byte n = grid.flat[8n+i] > thresh."""
    bits = (grid.reshape(-1, 8) > thresh).astype(np.uint8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)
