"""Timing of one joint training step (BASELINE configs[2]): LiDAR + camera rays through
NeRFNetwork.render with autograd, a simple L1/MSE loss, backward to all parameter gradients."""
import argparse, importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("selfsupervised-nvsf_b200")
S = importlib.import_module("selfsupervised-nvsf_b200.synth")
ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--steps", type=int, default=768)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--style", default="init")
ap.add_argument("--fused", type=int, default=1)
a = ap.parse_args()
torch.manual_seed(0)
m = pkg.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND, min_near=S.MIN_NEAR,
                    min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH).train()
if a.style == "trained":
    from oracle import field_init
    from oracle.field_oracle import FieldConfig
    cfg = FieldConfig(bound=S.BOUND, num_frames=S.NUM_FRAMES, time_resolution=S.TIME_RESOLUTION, min_near=S.MIN_NEAR,
                      min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)
    m.load_flat_params(field_init.make_params(cfg, seed=0, style="trained"))
if a.fused:
    m.fused_grad_accumulation = True
    for p in m.parameters():
        p.grad = torch.zeros_like(p)
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
lo, ld = map(dev, S.lidar_rays(a.rays, seed=1)); co, cd = map(dev, S.camera_rays(a.rays, seed=2))
t = torch.tensor([[0.4]], device="cuda")
gt_d = torch.rand(a.rays, device="cuda"); gt_i = torch.rand(a.rays, 2, device="cuda"); gt_c = torch.rand(a.rays, 3, device="cuda")
def step():
    ol = m.render(lo[None], ld[None], t, cal_lidar_color=True, staged=False, num_steps=a.steps, perturb=True)
    loss = (ol["depth_lidar"].view(-1) - gt_d).abs().mean() + ((ol["image_lidar"].view(-1, 2) - gt_i) ** 2).mean()
    loss.backward()
    oc = m.render(co[None], cd[None], t, cal_lidar_color=False, staged=False, num_steps=a.steps, perturb=True)
    loss2 = ((oc["image"].view(-1, 3) - gt_c) ** 2).mean()
    loss2.backward()
    return loss.detach() + loss2.detach()
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters): l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"train step: {a.rays}+{a.rays} rays x {a.steps}: {ms:.2f} ms/step, {2 * a.rays / ms * 1e3:.0f} rays/s, loss {l.item():.4f}")
gn = {n: float(p.grad.norm()) for n, p in m.named_parameters() if p.grad is not None}
print({k: f"{v:.3e}" for k, v in gn.items()})
assert all(np.isfinite(v) for v in gn.values())
