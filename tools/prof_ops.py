"""Profiling driver for the ray-marching operators (ncu target): one camera frame on the street-shell
grid through near_far -> march_rays_train -> composite_rays_train forward / backward, then a few
march_rays / composite_rays rounds of the inference loop.

    ncu --set full -k regex:"k_near_far|k_march|k_composite|k_packbits" python tools/prof_ops.py [kind] [fill] [sigma_scale]
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

S = cases.S
pkg = importlib.import_module("selfsupervised-nvsf_b200")
rm = pkg.raymarching
kind = sys.argv[1] if len(sys.argv) > 1 else "camera"
fill = sys.argv[2] if len(sys.argv) > 2 else "shell"
sig_scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
o, d, nears, fars, noises = cases.march_inputs(kind, -1, seed=1, perturb=True)
bf = cases.bitfield(fill, seed=2)
t_o, t_d, t_bf, t_no = map(dev, (o, d, bf, noises))
N = o.shape[0]
if kind == "camera":
    t_n, t_f = rm.near_far_from_aabb(t_o, t_d, dev(S.AABB), S.MIN_NEAR)
else:
    t_n, t_f = dev(nears), dev(fars)
grid = dev(S.density_grid(fill, seed=2))
rm.packbits(grid, cases.THRESH)
x, dd, dl, rays = rm.march_rays_train(t_o, t_d, S.BOUND, t_bf, S.CASCADE, S.GRID_SIZE, t_n, t_f, None, -1, False, -1,
                                      True, S.DT_GAMMA, 1024, t_no)
M = x.shape[0]
sig, rgb = cases.field_values(M, seed=3)
t_s = dev(sig * sig_scale).requires_grad_(True)
t_c = dev(rgb).requires_grad_(True)
ws, depth, image = rm.composite_rays_train(t_s, t_c, dl, rays, 1e-4)
(image.sum() + ws.sum()).backward()
# inference loop, two rounds
n_alive = N
rays_alive = torch.arange(N, dtype=torch.int32, device="cuda")
rays_t = t_n.clone()
wsum = torch.zeros(N, device="cuda"); dep = torch.zeros(N, device="cuda"); img = torch.zeros(N, 3, device="cuda")
for step in range(2):
    n_step = 16
    xs, ds, dls = rm.march_rays(n_alive, n_step, rays_alive, rays_t, t_o, t_d, S.BOUND, t_bf, S.CASCADE, S.GRID_SIZE,
                                t_n, t_f, -1, False, S.DT_GAMMA, 1024)
    sg, cl = cases.field_values(xs.shape[0], seed=4 + step)
    rm.composite_rays(n_alive, n_step, rays_alive, rays_t, dev(sg * sig_scale), dev(cl), dls, wsum, dep, img, 1e-2)
torch.cuda.synchronize()
print(kind, fill, "rays", N, "samples", M)
