"""TEST INFRASTRUCTURE — CPU restatement of the supervision terms of the reference's train_step.

Follows nvsf/nerf/trainer.py:188-219 (LiDAR: raydrop mask, label smoothing, the three weighted
criteria) and :503 (camera colour) line by line, with the criteria objects of
nvsf/scripts/main_nvsf.py:205-212 (torch.nn losses, reduction="none").  Pinned by construction: the
arithmetic is PyTorch's own loss modules, exactly the objects the reference instantiates.  Only
tests/ and __graft_entry__.smoke() may import this module."""
import torch


def loss_dict(scale=1.0):
    # main_nvsf.py:205-212
    return {
        "mse": torch.nn.MSELoss(reduction="none"),
        "l1": torch.nn.L1Loss(reduction="none"),
        "smoothl1": torch.nn.SmoothL1Loss(reduction="none", beta=0.1),
        "huber": torch.nn.HuberLoss(reduction="none", delta=0.2 * scale),
    }


def lidar_loss(depth_lidar, image_lidar, images_lidar, alpha_d=1.0, alpha_r=0.01, alpha_i=0.1, smooth_factor=0.0,
               depth_loss="l1", raydrop_loss="mse", intensity_loss="mse", scale=1.0):
    crit = loss_dict(scale)
    gt_raydrop = images_lidar[:, :, 0]                                  # trainer.py:188
    gt_intensity = images_lidar[:, :, 1] * gt_raydrop                   # :189
    gt_depth = images_lidar[:, :, 2] * gt_raydrop                       # :190
    pred_raydrop = image_lidar[:, :, 0]                                 # :206
    pred_intensity = image_lidar[:, :, 1] * gt_raydrop                  # :205
    pred_depth = depth_lidar * gt_raydrop                               # :206
    gt_raydrop_smooth = gt_raydrop.clamp(smooth_factor, 1 - smooth_factor)  # :211-213
    loss_d = alpha_d * crit[depth_loss](pred_depth, gt_depth)           # :219
    loss_rd = alpha_r * crit[raydrop_loss](pred_raydrop, gt_raydrop_smooth)  # :217
    loss_i = alpha_i * crit[intensity_loss](pred_intensity, gt_intensity)    # :218
    return loss_d + loss_rd + loss_i                                    # :219


def rgb_loss(pred_rgb, gt_rgb, alpha_rgb=1.0, rgb_loss="mse", scale=1.0):
    return alpha_rgb * loss_dict(scale)[rgb_loss](pred_rgb, gt_rgb)     # trainer.py:503
