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


# ---- the remaining terms of train_step, restated line by line (trainer.py line numbers on the right) ----
def flow_loss(flow_fn, cham_fn, pc, pc_forward=None, pc_backward=None):
    """trainer.py:237-265.  flow_fn(pc) -> {"flow_forward", "flow_backward"}; cham_fn(a, b) ->
    (dist1, dist2, idx1, idx2)."""
    pred_flow = flow_fn(pc)                                               # :244
    total = 0
    if pc_forward is not None:                                            # :249
        pc_pred = pc + pred_flow["flow_forward"]                          # :250
        dist1, dist2, _, _ = cham_fn(pc_pred.unsqueeze(0), pc_forward.unsqueeze(0))   # :253
        total = total + (dist1.sum() + dist2.sum()) * 0.5 + pred_flow["flow_forward"].abs().mean()   # :254-257
    if pc_backward is not None:                                           # :260
        pc_pred = pc + pred_flow["flow_backward"]
        dist1, dist2, _, _ = cham_fn(pc_pred.unsqueeze(0), pc_backward.unsqueeze(0))
        total = total + (dist1.sum() + dist2.sum()) * 0.5 + pred_flow["flow_backward"].abs().mean()
    return total


def los_loss(weights, z_vals, gt_depth, eps):
    """trainer.py:276-296 (gt_depth [B,N] already multiplied by the raydrop mask)."""
    import numpy as np
    g = gt_depth.reshape(z_vals.shape[0], 1)
    depth_mask = g > 0.0                                                  # :283
    mask_empty = (z_vals < (g - eps)) | (z_vals > (g + eps))              # :284
    loss_empty = ((mask_empty * weights) ** 2).sum() / depth_mask.sum()   # :285
    los = 0.1 * loss_empty                                                # :286
    mask_near = (z_vals > (g - eps)) & (z_vals < (g + eps))               # :288
    distance = mask_near * (z_vals - g)                                   # :289
    sigma = eps / 3.0                                                     # :290
    distr = 1.0 / (sigma * np.sqrt(2 * np.pi)) * torch.exp(-(distance ** 2 / (2 * sigma ** 2)))   # :291
    distr = distr / distr.max()                                           # :292
    distr = distr * mask_near                                             # :293
    loss_near = ((mask_near * weights - distr) ** 2).sum() / depth_mask.sum()   # :294
    return los + 0.1 * loss_near                                          # :295


_SOBEL_X = [[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]
_SOBEL_Y = [[-1, -2, -1], [0, 0, 0], [1, 2, 1]]


def _grads(d, sobel):
    import torch.nn.functional as F
    if sobel:                                                             # :318-331
        kx = torch.tensor(_SOBEL_X, dtype=torch.float32).unsqueeze(0).unsqueeze(0)
        ky = torch.tensor(_SOBEL_Y, dtype=torch.float32).unsqueeze(0).unsqueeze(0)
        return F.conv2d(d, kx, padding=1), F.conv2d(d, ky, padding=1)
    gx = d[:, :, :, :-1] - d[:, :, :, 1:]                                 # :333-336
    gx = torch.cat((gx, gx[:, :, :, -1:]), dim=3)
    gy = d[:, :, :-1, :] - d[:, :, 1:, :]
    gy = torch.cat((gy, gy[:, :, -1:, :]), dim=2)
    return gx, gy


def patch_grad_masks(pano_depth, rays_pano_inds, W_lidar, patch_h, patch_w, scale, thresh=0.05):
    """trainer.py:392-428."""
    patch_pxls_h = (rays_pano_inds // W_lidar).reshape(-1, patch_h, patch_w, 1).permute(0, 3, 1, 2).contiguous()
    patch_pxls_w = (rays_pano_inds % W_lidar).reshape(-1, patch_h, patch_w, 1).permute(0, 3, 1, 2).contiguous()
    num_patch = patch_pxls_h.shape[0]
    gx = (pano_depth[:, :-1] - pano_depth[:, 1:]) / scale
    gx = torch.cat((gx, gx[:, -1:]), dim=1)
    gy = (pano_depth[:-1, :] - pano_depth[1:, :]) / scale
    gy = torch.cat((gy, gy[-1:, :]), dim=0)
    gxx = gx[:, :-1].abs() - gx[:, 1:].abs()
    gxx = torch.cat((gxx, gxx[:, -1:]), dim=1)
    gyy = gy[:-1, :].abs() - gy[1:, :].abs()
    gyy = torch.cat((gyy, gyy[-1:, :]), dim=0)
    pxx = torch.gather(gxx.expand(num_patch, 1, -1, -1), 2, patch_pxls_h[:, :, :, :1].repeat(1, 1, 1, gxx.shape[1]))
    pxx = torch.gather(pxx, 3, patch_pxls_w)
    pyy = torch.gather(gyy.expand(num_patch, 1, -1, -1), 3, patch_pxls_w[:, :, :1, :].repeat(1, 1, gxx.shape[0], 1))
    pyy = torch.gather(pyy, 2, patch_pxls_h)
    return torch.where(pxx.abs() < thresh, 1, 0), torch.where(pyy.abs() < thresh, 1, 0)


def structural_loss(pred_depth, patch_h, patch_w, scale, gt_depth=None, gt_raydrop=None, grad_mask_x=None,
                    grad_mask_y=None, sobel_grad=False, grad_norm_smooth=False, spatial_smooth=False, tv_loss=False,
                    grad_loss=False, alpha_grad_norm=0.1, alpha_spatial=0.1, alpha_tv=0.1, alpha_grad=0.1,
                    depth_grad_loss="l1"):
    """trainer.py:306-462 on pred_depth [B,N] (already multiplied by the raydrop mask)."""
    crit = dict(loss_dict(scale), cos=torch.nn.CosineSimilarity())
    loss_sr = 0
    d = pred_depth.reshape(-1, patch_h, patch_w, 1).permute(0, 3, 1, 2).contiguous() / scale   # :307-310
    num_patch = d.shape[0]
    gx, gy = _grads(d, sobel_grad)
    if grad_norm_smooth:                                                  # :338-341
        loss_sr = loss_sr + alpha_grad_norm * (torch.exp(-gx.abs()) + torch.exp(-gy.abs()))
    if spatial_smooth:                                                    # :343-346
        loss_sr = loss_sr + alpha_spatial * (gx ** 2 + gy ** 2)
    if tv_loss:                                                           # :348-351
        loss_sr = loss_sr + alpha_tv * (gx.abs() + gy.abs())
    if grad_loss:                                                         # :354-462
        t = gt_depth.reshape(-1, patch_h, patch_w, 1).permute(0, 3, 1, 2).contiguous() / scale
        rd = gt_raydrop.reshape(-1, patch_h, patch_w, 1).permute(0, 3, 1, 2).contiguous()
        tx, ty = _grads(t, sobel_grad)
        mdx, mdy = rd * grad_mask_x, rd * grad_mask_y                     # :431-432
        if depth_grad_loss == "cos":                                      # :435-446
            lx = crit["cos"]((gx * mdx).reshape(num_patch, -1), (tx * mdx).reshape(num_patch, -1))
            ly = crit["cos"]((gy * mdy).reshape(num_patch, -1), (ty * mdy).reshape(num_patch, -1))
            lx = (1 - lx).reshape(num_patch, 1, 1, 1).expand(num_patch, 1, patch_h, patch_w)
            ly = (1 - ly).reshape(num_patch, 1, 1, 1).expand(num_patch, 1, patch_h, patch_w)
        else:                                                             # :448-449
            lx = crit[depth_grad_loss](gx * mdx, tx * mdx)
            ly = crit[depth_grad_loss](gy * mdy, ty * mdy)
        loss_sr = loss_sr + (alpha_grad * (lx + ly)).sum()                # :452-456
    return loss_sr
