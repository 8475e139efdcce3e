"""Loss head on the composited outputs — host-side mirror of the supervision terms of the reference's
`Trainer.train_step` (nvsf/nerf/trainer.py:188-219 LiDAR, :503 camera) with the element-wise
criteria of `main_nvsf.py:205-212`.  Same argument meaning as the reference options (`--alpha_d`,
`--alpha_r`, `--alpha_i`, `--alpha_rgb`, `--smooth_factor`, `--depth_loss`, `--raydrop_loss`,
`--intensity_loss`, `--rgb_loss`); the results are the un-reduced tensors the trainer sums
(`helper_loss(lidar_loss)`, trainer.py:545-547) and keeps for its error map (:551-560).

One CUDA kernel (csrc/loss.cu) computes the loss and its derivative with respect to the renderer's
outputs; `backward` only scales those buffers by the incoming gradient.  No CPU / PyTorch fallback:
`bce` and `cos` criteria are not built and raise."""
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
