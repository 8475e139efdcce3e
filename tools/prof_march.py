"""Profiling driver: one march_rays_train over a full camera frame (529 408 rays) on a chosen
occupancy fill.   ncu --set full -k regex:k_march_train python tools/prof_march.py camera full"""
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
fill = sys.argv[2] if len(sys.argv) > 2 else "full"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
o, d, nears, fars, noises = cases.march_inputs(kind, -1, seed=1, perturb=True)
bf = cases.bitfield(fill, seed=2)
t_o, t_d, t_bf, t_n, t_f, t_no = map(dev, (o, d, bf, nears, fars, noises))
for _ in range(reps):
    x, dd, l, r = rm.march_rays_train(t_o, t_d, S.BOUND, t_bf, S.CASCADE, S.GRID_SIZE, t_n, t_f, None, -1, False, -1,
                                      True, S.DT_GAMMA, 1024, t_no)
torch.cuda.synchronize()
print(kind, fill, "rays", o.shape[0], "samples", x.shape[0])
