"""TEST INFRASTRUCTURE.  numpy front-end of the C oracle (oracle/raymarching_oracle.c).

Function names and argument meaning follow the reference operator module
nvsf/nerf/raymarching/raymarching.py; inputs/outputs are numpy arrays on the host.
"""
import ctypes

import numpy as np

from . import build as _build

_lib = None
_P = ctypes.c_void_p
_U = ctypes.c_uint32
_F = ctypes.c_float


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build_oracle())
        for name in ("oracle_near_far_from_aabb", "oracle_sph_from_ray", "oracle_morton3D",
                     "oracle_morton3D_invert", "oracle_packbits", "oracle_march_rays_train",
                     "oracle_composite_rays_train_forward", "oracle_composite_rays_train_backward",
                     "oracle_march_rays", "oracle_composite_rays"):
            getattr(_lib, name).restype = None
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3), _f(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().oracle_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), _U(N), _F(min_near),
                                    _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().oracle_sph_from_ray(_p(rays_o), _p(rays_d), _F(radius), _U(N), _p(coords))
    return coords


def morton3D(coords):
    coords = _i(coords)
    N = coords.shape[0]
    out = np.empty(N, np.int32)
    lib().oracle_morton3D(_p(coords), _U(N), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i(indices)
    N = indices.shape[0]
    out = np.empty((N, 3), np.int32)
    lib().oracle_morton3D_invert(_p(indices), _U(N), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f(grid)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    lib().oracle_packbits(_p(grid), _U(N), _F(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises,
                     dt_gamma=0.0, max_steps=1024, M=None, counter=None):
    """Returns xyzs[M,3], dirs[M,3], deltas[M,2] (zero where unwritten), rays[N,3], counter[2]."""
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    nears, fars, noises = _f(nears), _f(fars), _f(noises)
    N = rays_o.shape[0]
    if M is None:
        M = N * max_steps
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    rays = np.zeros((N, 3), np.int32)
    counter = np.zeros(2, np.int32) if counter is None else _i(counter).copy()
    lib().oracle_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), _F(bound), _F(dt_gamma),
                                  _U(max_steps), _U(N), _U(C), _U(H), _U(M), _p(nears), _p(fars),
                                  _p(xyzs), _p(dirs), _p(deltas), _p(rays), _p(counter),
                                  _p(noises))
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f(sigmas), _f(rgbs), _f(deltas), _i(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    lib().oracle_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), _U(M),
                                              _U(N), _F(T_thresh), _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, weights_sum,
                                  image, T_thresh=1e-4):
    grad_ws, grad_image = _f(grad_ws), _f(grad_image)
    sigmas, rgbs, deltas, rays = _f(sigmas), _f(rgbs), _f(deltas), _i(rays)
    weights_sum, image = _f(weights_sum), _f(image)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gr = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().oracle_composite_rays_train_backward(_p(grad_ws), _p(grad_image), _p(sigmas), _p(rgbs),
                                               _p(deltas), _p(rays), _p(weights_sum), _p(image),
                                               _U(M), _U(N), _F(T_thresh), _p(gs), _p(gr))
    return gs, gr


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H,
               nears, fars, noises, dt_gamma=0.0, max_steps=1024, M=None):
    rays_alive, rays_t = _i(rays_alive), _f(rays_t)
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    nears, fars, noises = _f(nears), _f(fars), _f(noises)
    if M is None:
        M = n_alive * n_step
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    lib().oracle_march_rays(_U(n_alive), _U(n_step), _p(rays_alive), _p(rays_t), _p(rays_o),
                            _p(rays_d), _F(bound), _F(dt_gamma), _U(max_steps), _U(C), _U(H),
                            _p(bitfield), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas),
                            _p(noises))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth,
                   image, T_thresh=1e-2):
    """Returns updated copies (rays_alive, rays_t, weights_sum, depth, image)."""
    rays_alive, rays_t = _i(rays_alive).copy(), _f(rays_t).copy()
    sigmas, rgbs, deltas = _f(sigmas), _f(rgbs), _f(deltas)
    weights_sum, depth, image = _f(weights_sum).copy(), _f(depth).copy(), _f(image).copy()
    lib().oracle_composite_rays(_U(n_alive), _U(n_step), _F(T_thresh), _p(rays_alive), _p(rays_t),
                                _p(sigmas), _p(rgbs), _p(deltas), _p(weights_sum), _p(depth),
                                _p(image))
    return rays_alive, rays_t, weights_sum, depth, image
