"""Loss head on the composited outputs — host-side mirror of the supervision terms of the reference's
`Trainer.train_step` (nvsf/nerf/trainer.py:188-219 LiDAR, :503 camera) with the element-wise
criteria of `main_nvsf.py:205-212`.  Same argument meaning as the reference options (`--alpha_d`,
`--alpha_r`, `--alpha_i`, `--alpha_rgb`, `--smooth_factor`, `--depth_loss`, `--raydrop_loss`,
`--intensity_loss`, `--rgb_loss`); the results are the un-reduced tensors the trainer sums
(`helper_loss(lidar_loss)`, trainer.py:545-547) and keeps for its error map (:551-560).

One CUDA kernel (csrc/loss.cu) computes the loss and its derivative with respect to the renderer's
outputs; `backward` only scales those buffers by the incoming gradient.  No CPU / PyTorch fallback.
Further down: the scene-flow loss (trainer.py:237-265), the URF line-of-sight loss (:276-296) and the
structural regularisation of depth patches (:297-462); the `bce` criterion is not built and raises."""
import ctypes

import torch

from ._lib import NvsfError, check, lib, ptr, stream_ptr

_KINDS = {"l1": 0, "mse": 1, "smoothl1": 2, "huber": 3}


class LidarLossCfg(ctypes.Structure):
    """struct nvsf_lidar_loss_cfg (include/nvsf_b200.h Part 5)."""
    _fields_ = [("alpha_d", ctypes.c_float), ("alpha_r", ctypes.c_float), ("alpha_i", ctypes.c_float),
                ("smooth", ctypes.c_float), ("depth_kind", ctypes.c_int32), ("raydrop_kind", ctypes.c_int32),
                ("intensity_kind", ctypes.c_int32), ("depth_param", ctypes.c_float),
                ("raydrop_param", ctypes.c_float), ("intensity_param", ctypes.c_float)]


def _kind(name, scale):
    """criterion name -> (kind, parameter): SmoothL1Loss(beta=0.1), HuberLoss(delta=0.2*scale)
    (main_nvsf.py:208-209)."""
    if name not in _KINDS:
        raise NvsfError(f"criterion {name!r} is not built on the GPU path (have {sorted(_KINDS)})")
    return _KINDS[name], {"smoothl1": 0.1, "huber": 0.2 * scale}.get(name, 0.0)


def _f32c(t):
    return t.detach().to(dtype=torch.float32).contiguous()


class _LidarLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, image, gt, cfg):
        if not depth.is_cuda:
            raise NvsfError("lidar_loss: tensors must live on the GPU (no CPU path)")
        d, im, g = _f32c(depth).view(-1), _f32c(image).view(-1, 2), _f32c(gt).view(-1, 3)
        n = d.shape[0]
        if im.shape[0] != n or g.shape[0] != n:
            raise NvsfError("lidar_loss: depth [.,N], image [.,N,2] and images_lidar [.,N,3] disagree")
        loss = torch.empty(n, dtype=torch.float32, device=d.device)
        g_depth = torch.empty(n, dtype=torch.float32, device=d.device)
        g_image = torch.empty(n, 2, dtype=torch.float32, device=d.device)
        check(lib().nvsf_loss_lidar(ptr(d), ptr(im), ptr(g), n, ctypes.byref(cfg), ptr(loss), ptr(g_depth),
                                    ptr(g_image), stream_ptr()), "loss_lidar")
        ctx.save_for_backward(g_depth, g_image)
        ctx.shapes = (depth.shape, image.shape)
        return loss.view(depth.shape)

    @staticmethod
    def backward(ctx, g_loss):
        g_depth, g_image = ctx.saved_tensors
        gl = g_loss.reshape(-1)
        return (g_depth * gl).view(ctx.shapes[0]), (g_image * gl[:, None]).view(ctx.shapes[1]), None, None


def lidar_loss(depth_lidar, image_lidar, images_lidar, alpha_d=1.0, alpha_r=0.01, alpha_i=0.1, smooth_factor=0.0,
               depth_loss="l1", raydrop_loss="mse", intensity_loss="mse", scale=1.0):
    """`lidar_loss` [B, N] of trainer.py:188-219 from `outputs_lidar["depth_lidar"]` [B, N],
    `outputs_lidar["image_lidar"]` [B, N, 2] (raydrop, intensity) and the ground truth `images_lidar`
    [B, N, 3] (raydrop mask, intensity, depth)."""
    cfg = LidarLossCfg()
    cfg.alpha_d, cfg.alpha_r, cfg.alpha_i, cfg.smooth = alpha_d, alpha_r, alpha_i, smooth_factor
    cfg.depth_kind, cfg.depth_param = _kind(depth_loss, scale)
    cfg.raydrop_kind, cfg.raydrop_param = _kind(raydrop_loss, scale)
    cfg.intensity_kind, cfg.intensity_param = _kind(intensity_loss, scale)
    return _LidarLoss.apply(depth_lidar, image_lidar, images_lidar, cfg)


class _ElemLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, kind, param, alpha):
        if not pred.is_cuda:
            raise NvsfError("rgb_loss: tensors must live on the GPU (no CPU path)")
        p, g = _f32c(pred).view(-1), _f32c(gt).view(-1)
        if p.shape != g.shape:
            raise NvsfError("rgb_loss: prediction and ground truth disagree in size")
        loss, g_pred = torch.empty_like(p), torch.empty_like(p)
        check(lib().nvsf_loss_elementwise(ptr(p), ptr(g), p.numel(), kind, param, alpha, ptr(loss), ptr(g_pred),
                                          stream_ptr()), "loss_elementwise")
        ctx.save_for_backward(g_pred)
        ctx.shape = pred.shape
        return loss.view(pred.shape)

    @staticmethod
    def backward(ctx, g_loss):
        (g_pred,) = ctx.saved_tensors
        return (g_pred * g_loss.reshape(-1)).view(ctx.shape), None, None, None, None


def rgb_loss(pred_rgb, gt_rgb, alpha_rgb=1.0, rgb_loss="mse", scale=1.0):
    """`alpha_rgb * criterion["rgb"](pred_rgb, gt_rgb)` [B, N, 3] of trainer.py:503."""
    kind, param = _kind(rgb_loss, scale)
    return _ElemLoss.apply(pred_rgb, gt_rgb, kind, float(param), float(alpha_rgb))


# ---------------------------------------------------------------------------------------------------
# the remaining loss terms of train_step: scene-flow (Chamfer on warped clouds), URF line-of-sight,
# structural regularisation of depth patches
# ---------------------------------------------------------------------------------------------------
def flow_loss(model, pc, time, pc_forward=None, pc_backward=None, chamfer=None):
    """Scene-flow loss of trainer.py:237-265: `pc` [M,3] is the LiDAR cloud of the current frame
    (normalised coordinates), `pc_forward` / `pc_backward` those of the next / previous frame (None when
    the frame has no such neighbour).  Per direction: 0.5 (sum dist1 + sum dist2) of the Chamfer distance
    between pc + flow and the neighbour cloud, plus mean |flow|.  Differentiable with respect to
    flow_net (NeRFNetwork.flow, csrc/train.cu) through the Chamfer backward (csrc/chamfer.cu)."""
    from .chamfer import chamfer_3DDist
    cham = chamfer if chamfer is not None else chamfer_3DDist()
    pc = pc.detach().to(dtype=torch.float32).contiguous()
    f = model.flow(pc, time)
    total = 0
    for key, other in (("flow_forward", pc_forward), ("flow_backward", pc_backward)):
        if other is None:
            continue
        pred = pc + f[key]
        d1, d2, _, _ = cham(pred.unsqueeze(0), other.detach().to(dtype=torch.float32).contiguous().unsqueeze(0))
        total = total + (d1.sum() + d2.sum()) * 0.5 + f[key].abs().mean()
    return total


class _LosLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, z_vals, gt_depth, eps):
        if not weights.is_cuda:
            raise NvsfError("los_loss: tensors must live on the GPU (no CPU path)")
        w, z, g = _f32c(weights), _f32c(z_vals), _f32c(gt_depth).view(-1)
        R, T = w.shape
        if z.shape != w.shape or g.shape[0] != R:
            raise NvsfError("los_loss: weights [R,T], z_vals [R,T] and gt_depth [R] disagree")
        loss = torch.zeros(1, dtype=torch.float32, device=w.device)
        g_w = torch.empty_like(w)
        ws = torch.empty(16, dtype=torch.uint8, device=w.device)
        check(lib().nvsf_loss_los(ptr(w), ptr(z), ptr(g), R, T, float(eps), ptr(loss), ptr(g_w), ptr(ws), 16,
                                  stream_ptr()), "loss_los")
        ctx.save_for_backward(g_w)
        ctx.shape = weights.shape
        return loss.view(())

    @staticmethod
    def backward(ctx, g_loss):
        (g_w,) = ctx.saved_tensors
        return (g_w * g_loss).view(ctx.shape), None, None, None


def urf_eps(global_step, iters):
    """eps of trainer.py:278."""
    return 0.02 * 0.1 ** min(global_step / iters, 1)


def los_loss(weights, z_vals, gt_depth, eps):
    """Line-of-sight loss of Urban Radiance Fields as train_step applies it (trainer.py:276-296) to
    `outputs_lidar["weights"]` [B*N,T] and `["z_vals"]` with gt_depth [B,N] (already multiplied by the
    raydrop mask); returns the scalar `los_loss`."""
    return _LosLoss.apply(weights, z_vals, gt_depth, float(eps))


class PatchLossCfg(ctypes.Structure):
    """struct nvsf_patch_loss_cfg (include/nvsf_b200.h Part 5)."""
    _fields_ = [("scale", ctypes.c_float), ("sobel", ctypes.c_int32), ("grad_norm_smooth", ctypes.c_int32),
                ("spatial_smooth", ctypes.c_int32), ("tv_loss", ctypes.c_int32), ("grad_loss", ctypes.c_int32),
                ("alpha_grad_norm", ctypes.c_float), ("alpha_spatial", ctypes.c_float), ("alpha_tv", ctypes.c_float),
                ("alpha_grad", ctypes.c_float), ("grad_kind", ctypes.c_int32), ("grad_param", ctypes.c_float)]


def patch_grad_masks(pano_depth, rays_pano_inds, patch_h, patch_w, scale, thresh=0.05):
    """grad_mask_x / grad_mask_y [P,1,h,w] of trainer.py:392-428 from the range image
    `data['pano_frame'][0, ..., 2]` [H,W] and `data['rays_pano_inds']` (P*h*w pixel ids)."""
    if not pano_depth.is_cuda:
        raise NvsfError("patch_grad_masks: tensors must live on the GPU (no CPU path)")
    d = _f32c(pano_depth)
    H, W = d.shape
    inds = rays_pano_inds.detach().to(device=d.device, dtype=torch.int64).contiguous().view(-1)
    P = inds.numel() // (patch_h * patch_w)
    mx = torch.empty(P, 1, patch_h, patch_w, dtype=torch.float32, device=d.device)
    my = torch.empty_like(mx)
    check(lib().nvsf_patch_grad_masks(ptr(d), H, W, ptr(inds), P, patch_h, patch_w, float(scale), float(thresh),
                                      ptr(mx), ptr(my), stream_ptr()), "patch_grad_masks")
    return mx, my


class _PatchLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt_depth, gt_raydrop, mask_x, mask_y, h, w, cfg):
        if not pred.is_cuda:
            raise NvsfError("structural_loss: tensors must live on the GPU (no CPU path)")
        p = _f32c(pred).view(-1, h, w)
        P = p.shape[0]
        opt = lambda t: None if t is None else _f32c(t).view(-1, h, w)
        gd, gr, mx, my = opt(gt_depth), opt(gt_raydrop), opt(mask_x), opt(mask_y)
        loss_map = torch.empty(P, 1, h, w, dtype=torch.float32, device=p.device)
        grad_loss = torch.empty(P, dtype=torch.float32, device=p.device)
        check(lib().nvsf_loss_patch(ptr(p), ptr(gd), ptr(gr), ptr(mx), ptr(my), P, h, w, ctypes.byref(cfg),
                                    ptr(loss_map), ptr(grad_loss), None, None, None, stream_ptr()), "loss_patch")
        ctx.tensors = (p, gd, gr, mx, my)
        ctx.cfg, ctx.shape, ctx.hw = cfg, pred.shape, (P, h, w)
        return loss_map, grad_loss

    @staticmethod
    def backward(ctx, g_map, g_grad):
        p, gd, gr, mx, my = ctx.tensors
        P, h, w = ctx.hw
        gm = None if g_map is None else _f32c(g_map).view(P, h, w)
        gg = None if g_grad is None else _f32c(g_grad).view(P)
        g_pred = torch.empty_like(p)
        check(lib().nvsf_loss_patch(ptr(p), ptr(gd), ptr(gr), ptr(mx), ptr(my), P, h, w, ctypes.byref(ctx.cfg),
                                    None, None, ptr(gm), ptr(gg), ptr(g_pred), stream_ptr()), "loss_patch(backward)")
        return g_pred.view(ctx.shape), None, None, None, None, None, None, None


def structural_loss(pred_depth, patch_h, patch_w, scale, gt_depth=None, gt_raydrop=None, grad_mask_x=None,
                    grad_mask_y=None, sobel_grad=False, grad_norm_smooth=False, spatial_smooth=False, tv_loss=False,
                    grad_loss=False, alpha_grad_norm=0.1, alpha_spatial=0.1, alpha_tv=0.1, alpha_grad=0.1,
                    depth_grad_loss="l1"):
    """Structural regularisation of train_step (trainer.py:297-462) on `pred_depth` [B,N] (depth_lidar times
    the raydrop mask; N = P*h*w rays sampled as P patches).  Returns `loss_sr` as the reference forms it:
    the [P,1,h,w] map of the element-wise terms (grad_norm / spatial / tv, when any is enabled) plus the
    scalar `grad_loss.sum()` (when grad_loss is enabled)."""
    cfg = PatchLossCfg()
    cfg.scale, cfg.sobel = float(scale), int(bool(sobel_grad))
    cfg.grad_norm_smooth, cfg.spatial_smooth, cfg.tv_loss = int(bool(grad_norm_smooth)), int(bool(spatial_smooth)), int(bool(tv_loss))
    cfg.grad_loss = int(bool(grad_loss))
    cfg.alpha_grad_norm, cfg.alpha_spatial, cfg.alpha_tv, cfg.alpha_grad = alpha_grad_norm, alpha_spatial, alpha_tv, alpha_grad
    if depth_grad_loss == "cos":
        cfg.grad_kind, cfg.grad_param = 4, 0.0
    else:
        cfg.grad_kind, cfg.grad_param = _kind(depth_grad_loss, scale)
    loss_map, gl = _PatchLoss.apply(pred_depth, gt_depth, gt_raydrop, grad_mask_x, grad_mask_y, int(patch_h),
                                    int(patch_w), cfg)
    loss_sr = 0
    if grad_norm_smooth or spatial_smooth or tv_loss:
        loss_sr = loss_sr + loss_map
    if grad_loss:
        loss_sr = loss_sr + gl.sum()
    return loss_sr
