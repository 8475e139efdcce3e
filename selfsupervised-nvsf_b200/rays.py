"""Ray generation on the device: the reference's ``get_lidar_rays`` / ``get_rays``
(nvsf/nerf/dataset/dataset_utils.py:369-536, 539-687) with the same arguments and result dict.

Pixel selection (which ids to render) stays a torch RNG call exactly like the reference's
(`torch.randint(0, H*W, [N])`, so a seeded run picks the same pixels); turning pixel ids and a pose
into origins and directions is one kernel launch per pose (`nvsf_get_lidar_rays` /
`nvsf_get_rays`) instead of ~20 ATen launches over [B, H*W] meshgrids.
"""
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _select(H, W, N, patch_size, device, use_error_map):
    """Pixel ids [n] (int64) or None for the full image in row-major order."""
    if use_error_map:
        raise NotImplementedError("error-map guided sampling is a trainer feature outside the render hot path")
    if N <= 0:
        return None
    N = min(N, H * W)
    if isinstance(patch_size, int):
        ph, pw = patch_size, patch_size
    elif len(patch_size) == 1:
        ph, pw = patch_size[0], patch_size[0]
    else:
        ph, pw = patch_size
    if ph > 1:  # random patches, top-left corners may repeat (dataset_utils.py:444-465)
        num_patch = N // (ph * pw)
        inds_x = torch.randint(0, W - pw, size=[num_patch], device=device)
        inds_y = torch.randint(0, H - ph, size=[num_patch], device=device)
        oy, ox = torch.meshgrid(torch.arange(ph, device=device), torch.arange(pw, device=device), indexing="ij")
        rows = inds_y[:, None] + oy.reshape(1, -1)
        cols = inds_x[:, None] + ox.reshape(1, -1)
        return (rows * W + cols).reshape(-1)
    return torch.randint(0, H * W, size=[N], device=device)


def _generate(kind, poses, scalars, H, W, N, patch_size, use_error_map):
    L = _lib.lib()
    if not poses.is_cuda:
        poses = poses.cuda()
    poses = poses.detach().to(torch.float32).contiguous()
    dev = poses.device
    B = poses.shape[0]
    inds = _select(H, W, N, patch_size, dev, use_error_map)
    n = H * W if inds is None else inds.shape[0]
    rays_o = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
    fn = L.nvsf_get_lidar_rays if kind == "lidar" else L.nvsf_get_rays
    for b in range(B):
        check(fn(ptr(poses[b]), ptr(inds), n, H, W, *scalars, ptr(rays_o[b]), ptr(rays_d[b]), stream_ptr()),
              f"get_{kind}_rays")
    if inds is None:
        inds = torch.arange(H * W, device=dev)
    return {"rays_o": rays_o, "rays_d": rays_d, "inds": inds.expand([B, n])}


@torch.no_grad()
def get_lidar_rays(poses, intrinsics, intrinsics_hoz, H, W, N=-1, patch_size=1, error_map=None,
                   use_error_map=False):
    """poses [B,4,4] lidar2world; intrinsics = (fov_up, fov) and intrinsics_hoz = (fov_up, fov) in
    degrees -> {'rays_o' [B,n,3], 'rays_d' [B,n,3], 'inds' [B,n]}."""
    fov_up, fov = float(intrinsics[0]), float(intrinsics[1])
    fov_hoz = float(intrinsics_hoz[1])
    return _generate("lidar", poses, (fov_up, fov, fov_hoz), H, W, N, patch_size, use_error_map)


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, patch_size=1, error_map=None, use_error_map=False):
    """poses [B,4,4] cam2world; intrinsics = 3x3 (or 3x4) pinhole matrix."""
    fx, fy = float(intrinsics[0][0]), float(intrinsics[1][1])
    cx, cy = float(intrinsics[0][2]), float(intrinsics[1][2])
    return _generate("camera", poses, (fx, fy, cx, cy), H, W, N, patch_size, use_error_map)
