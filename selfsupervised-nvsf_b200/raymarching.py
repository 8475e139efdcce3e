"""Drop-in for the reference operator module ``nvsf.nerf.raymarching.raymarching``.

Same nine callables, same positional/keyword parameters, defaults, output shapes
and dtypes, same in-place conventions (reference
nvsf/nerf/raymarching/raymarching.py:18,54,87,113,139,174-191,295,370-388,466-479).
Every op runs a hand-written sm_100a kernel through the C ABI in
include/nvsf_b200.h; there is no CPU or PyTorch fallback.

To swap it into the reference::

    import sys, nvsf_b200
    sys.modules["nvsf.nerf.raymarching.raymarching"] = nvsf_b200.raymarching

Behavioural notes (all within what the reference leaves unspecified):
  * `march_rays_train` returns `rays` rows in ray-id order with prefix-sum
    offsets (the reference's order depends on atomicAdd scheduling);
  * kernels are launched on the current torch stream, not the legacy default
    stream;
  * invalid arguments raise instead of producing asynchronous CUDA errors.
"""
import torch
from torch.autograd import Function

from . import _lib
from ._lib import check, ptr, stream_ptr

_f32 = torch.float32
last_step_counter = None


def _fwd(fn):
    return torch.amp.custom_fwd(_lib.device_guard(fn), device_type="cuda", cast_inputs=_f32)


def _bwd(fn):
    return torch.amp.custom_bwd(_lib.device_guard(fn), device_type="cuda")


def _cuda(t):
    return t if t.is_cuda else t.cuda()


def _rays(t):
    return _cuda(t).to(_f32).contiguous().view(-1, 3)


def _pad(m, align):
    # the reference adds a full `align` when m is already aligned (raymarching.py:231-232,278-279)
    return m + (align - m % align) if align > 0 else m


# ----------------------------------------------------------------------------
# utils
# ----------------------------------------------------------------------------
class _near_far_from_aabb(Function):
    @staticmethod
    @_fwd
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        """rays_o/rays_d [N,3], aabb [6] -> nears [N], fars [N] (reference raymarching.py:15-45)."""
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        aabb = aabb.to(device=rays_o.device, dtype=_f32).contiguous()
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=_f32, device=rays_o.device)
        fars = torch.empty(N, dtype=_f32, device=rays_o.device)
        check(_lib.lib().nvsf_near_far_from_aabb(ptr(rays_o), ptr(rays_d), ptr(aabb), N,
                                                 float(min_near), ptr(nears), ptr(fars),
                                                 stream_ptr()), "near_far_from_aabb")
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @_fwd
    def forward(ctx, rays_o, rays_d, radius):
        """Background-sphere (theta, phi) in [-1,1]^2, [N,2] (reference raymarching.py:51-79)."""
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=_f32, device=rays_o.device)
        check(_lib.lib().nvsf_sph_from_ray(ptr(rays_o), ptr(rays_d), float(radius), N,
                                           ptr(coords), stream_ptr()), "sph_from_ray")
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        """coords [N,3] int32 -> Morton indices [N] int32 (reference raymarching.py:85-105)."""
        coords = _cuda(coords).int().contiguous()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        check(_lib.lib().nvsf_morton3D(ptr(coords), N, ptr(indices), stream_ptr()), "morton3D")
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        """indices [N] int32 -> coords [N,3] int32 (reference raymarching.py:111-130)."""
        indices = _cuda(indices).int().contiguous()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        check(_lib.lib().nvsf_morton3D_invert(ptr(indices), N, ptr(coords), stream_ptr()),
              "morton3D_invert")
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @_fwd
    def forward(ctx, grid, thresh, bitfield=None):
        """grid [C, H^3] f32 -> bitfield [C*H^3/8] u8 (reference raymarching.py:136-161)."""
        grid = _cuda(grid).contiguous()
        N = grid.shape[0] * grid.shape[1] // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        check(_lib.lib().nvsf_packbits(ptr(grid), N, float(thresh), ptr(bitfield), stream_ptr()),
              "packbits")
        return bitfield


packbits = _packbits.apply


# ----------------------------------------------------------------------------
# train
# ----------------------------------------------------------------------------
class _march_rays_train(Function):
    @staticmethod
    @_fwd
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars,
                step_counter=None, mean_count=-1, perturb=False, align=-1,
                force_all_rays=False, dt_gamma=0, max_steps=1024, noises=None):
        """Occupancy-grid ray marching for training (reference raymarching.py:171-286).

        Returns xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3] (id, offset, count).
        `noises` (extra, optional) lets a caller supply the per-ray jitter instead of
        torch.rand — used by the parity tests.
        """
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        dev = rays_o.device
        density_bitfield = _cuda(density_bitfield).contiguous()
        nears = _cuda(nears).contiguous()
        fars = _cuda(fars).contiguous()
        N = rays_o.shape[0]
        L = _lib.lib()

        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        # a caller-owned counter may be non-zero on entry (it is accumulated, never reset:
        # raymarching.py:242-245); rows below its start value are then nobody's and must read 0.
        alloc = torch.empty if step_counter is None else torch.zeros
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        global last_step_counter
        last_step_counter = step_counter   # device-side (samples, rays) of the latest call, for callers that size
                                           # M from `mean_count` and check the count later without a sync here
        if noises is None:
            noises = (torch.rand(N, dtype=_f32, device=dev) if perturb
                      else torch.zeros(N, dtype=_f32, device=dev))
        else:
            noises = _cuda(noises).to(_f32).contiguous()
        ws_bytes = L.nvsf_march_rays_train_workspace_bytes(N)
        ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=dev)
        st = stream_ptr()
        geom = (ptr(rays_o), ptr(rays_d), ptr(density_bitfield), float(bound), float(dt_gamma),
                int(max_steps), N, int(C), int(H))

        # phase 1: counts + prefix sums -> rays, step_counter
        check(L.nvsf_march_rays_train_count(*geom, ptr(nears), ptr(fars), ptr(rays),
                                            ptr(step_counter), ptr(noises), ptr(ws), ws_bytes,
                                            st), "march_rays_train(count)")
        if force_all_rays or mean_count <= 0:
            # the reference sizes N*max_steps rows and slices after this same D2H read
            # (raymarching.py:276-282); here the read happens first so only M rows exist.
            M = _pad(int(step_counter[0].item()), align)
        else:
            M = _pad(int(mean_count), align)
        xyzs = alloc(M, 3, dtype=_f32, device=dev)
        dirs = alloc(M, 3, dtype=_f32, device=dev)
        deltas = alloc(M, 2, dtype=_f32, device=dev)
        # phase 2: emit samples; rows not covered by a ray are zero-filled by the kernel
        # (the workspace still holds the first samples of every ray from phase 1)
        check(L.nvsf_march_rays_train_write_ws(*geom, M, ptr(nears), ptr(fars), ptr(xyzs),
                                               ptr(dirs), ptr(deltas), ptr(rays),
                                               ptr(step_counter), ptr(noises), M, ptr(ws),
                                               ws_bytes, st),
              "march_rays_train(write)")
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @_fwd
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        """sigmas [M], rgbs [M,3], deltas [M,2], rays [N,3] -> weights_sum [N], depth [N],
        image [N,3] (reference raymarching.py:292-325)."""
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        deltas = deltas.contiguous()
        rays = rays.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=_f32, device=dev)
        depth = torch.empty(N, dtype=_f32, device=dev)
        image = torch.empty(N, 3, dtype=_f32, device=dev)
        check(_lib.lib().nvsf_composite_rays_train_forward(
            ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays), M, N, float(T_thresh),
            ptr(weights_sum), ptr(depth), ptr(image), stream_ptr()), "composite_rays_train fwd")
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh]
        return weights_sum, depth, image

    @staticmethod
    @_bwd
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is ignored, as in the reference (raymarching.py:330)
        grad_weights_sum = grad_weights_sum.contiguous()
        grad_image = grad_image.contiguous()
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        check(_lib.lib().nvsf_composite_rays_train_backward(
            ptr(grad_weights_sum), ptr(grad_image), ptr(sigmas), ptr(rgbs), ptr(deltas),
            ptr(rays), ptr(weights_sum), ptr(image), M, N, float(T_thresh), ptr(grad_sigmas),
            ptr(grad_rgbs), stream_ptr()), "composite_rays_train bwd")
        return grad_sigmas, grad_rgbs, None, None, None


composite_rays_train = _composite_rays_train.apply


# ----------------------------------------------------------------------------
# infer
# ----------------------------------------------------------------------------
class _march_rays(Function):
    @staticmethod
    @_fwd
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound,
                density_bitfield, C, H, near, far, align=-1, perturb=False, dt_gamma=0,
                max_steps=1024, noises=None):
        """Inference marcher (reference raymarching.py:367-457): up to n_step samples per alive
        ray from rays_t; returns xyzs/dirs [M,3], deltas [M,2] with M = n_alive*n_step (+align)."""
        rays_o, rays_d = _rays(rays_o), _rays(rays_d)
        dev = rays_o.device
        density_bitfield = _cuda(density_bitfield).contiguous()
        M = _pad(n_alive * n_step, align)
        xyzs = torch.empty(M, 3, dtype=_f32, device=dev)
        dirs = torch.empty(M, 3, dtype=_f32, device=dev)
        deltas = torch.empty(M, 2, dtype=_f32, device=dev)
        if noises is None:
            noises = (torch.rand(n_alive, dtype=_f32, device=dev) if perturb
                      else torch.zeros(n_alive, dtype=_f32, device=dev))
        else:
            noises = _cuda(noises).to(_f32).contiguous()
        check(_lib.lib().nvsf_march_rays(
            int(n_alive), int(n_step), ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d),
            float(bound), float(dt_gamma), int(max_steps), int(C), int(H),
            ptr(density_bitfield), ptr(near), ptr(far), ptr(xyzs), ptr(dirs), ptr(deltas),
            ptr(noises), M, stream_ptr()), "march_rays")
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @_fwd
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum,
                depth, image, T_thresh=1e-2):
        """In-place inference compositing (reference raymarching.py:463-507); returns ()."""
        check(_lib.lib().nvsf_composite_rays(
            int(n_alive), int(n_step), float(T_thresh), ptr(rays_alive), ptr(rays_t),
            ptr(sigmas.contiguous()), ptr(rgbs.contiguous()), ptr(deltas.contiguous()),
            ptr(weights_sum), ptr(depth), ptr(image), stream_ptr()), "composite_rays")
        return tuple()


composite_rays = _composite_rays.apply

__all__ = [
    "near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits",
    "march_rays_train", "composite_rays_train", "march_rays", "composite_rays",
]
