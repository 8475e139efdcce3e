"""Micro-benchmark of the occupancy-skipping renders (BASELINE configs[3]: camera 376x1408 RGB with
occupancy-grid skipping; also the LiDAR frame): run_cuda one-shot vs alive-ray loop, per-stage
CUDA-event times (march / density / heads / composite), update_extra_state and ray generation.
Development tool; bench.py is the judged benchmark.

    python tools/bench_march_render.py [--fills shell,random5,full] [--iters 5]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("selfsupervised-nvsf_b200")
S = importlib.import_module("selfsupervised-nvsf_b200.synth")
rm = pkg.raymarching


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        a, b = ev(), ev()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fills", default="shell,random5")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--style", default="default")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    m = pkg.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                        min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH).eval()
    out = []
    # ray generation (full frames)
    R, t = S.random_pose(0)
    P = np.eye(4, dtype=np.float32)
    P[:3, :3], P[:3, 3] = R, t
    P = torch.from_numpy(P).cuda()[None]
    K = np.array([[552.554261, 0, 682.049453], [0, 552.554261, 238.769549], [0, 0, 1]], np.float32)
    ms = timed(lambda: pkg.rays.get_rays(P, K, S.CAM_H, S.CAM_W, -1), args.iters)
    out.append(dict(op="get_rays", N=S.CAM_H * S.CAM_W, ms=ms, GBps=S.CAM_H * S.CAM_W * 24 / ms / 1e6))
    ms = timed(lambda: pkg.rays.get_lidar_rays(P, [2.0, 26.9], [180.0, 360.0], S.LIDAR_H, S.LIDAR_W, -1), args.iters)
    out.append(dict(op="get_lidar_rays", N=S.LIDAR_H * S.LIDAR_W, ms=ms))
    # occupancy-grid update (2 x 128^3 cells, one frame time)
    for lidar in (True, False):
        ms = timed(lambda: m.update_extra_state(0.5, cal_lidar_color=lidar), args.iters)
        occ = float(np.unpackbits(m.density_bitfield(lidar).cpu().numpy()).mean())
        out.append(dict(op="update_extra_state", lidar=lidar, cells=2 * 128 ** 3, ms=ms, occupancy=occ,
                        mean_density=float(m.mean_density(lidar))))
    for kind in ("camera", "lidar"):
        lidar = kind == "lidar"
        o, d = (S.lidar_rays if lidar else S.camera_rays)(-1, seed=0)
        to, td = torch.from_numpy(o).cuda()[None], torch.from_numpy(d).cuda()[None]
        N = o.shape[0]
        for fill in args.fills.split(",") + ["own"]:
            if fill == "own":
                bf = m.density_bitfield(lidar)
            else:
                bf = torch.from_numpy(S.packbits_np(S.density_grid(fill), 0.01)).cuda()
            for one_shot in (True, False):
                kw = dict(cal_lidar_color=lidar, dt_gamma=S.DT_GAMMA, T_thresh=1e-2, density_bitfield=bf,
                          one_shot=one_shot)
                ms = timed(lambda: m.run_cuda(to, td, 0.5, **kw), args.iters)
                rec = dict(op="run_cuda", kind=kind, fill=fill, one_shot=one_shot, N=N, ms=ms,
                           samples=int(m.last_run_cuda_samples), rays_per_s=N / ms * 1e3)
                out.append(rec)
            # stage split of the one-shot path
            oo, dd, nears, fars, _ = m._rays_setup(to, td, lidar, 1, False, None)
            m.prepare(0.5, lidar)
            res = {}
            def march():
                res["m"] = rm.march_rays_train(oo, dd, S.BOUND, bf, S.CASCADE, S.GRID_SIZE, nears, fars, None, -1,
                                               False, -1, True, S.DT_GAMMA, 1024, None)
            t_march = timed(march, args.iters)
            xyzs, dirs, deltas, rays = res["m"]
            def dens():
                res["d"] = m._density_raw(xyzs, 0.5, lidar)
            t_den = timed(dens, args.iters)
            sigma, geo16 = res["d"][0], res["d"][1]
            def heads():
                res["c"] = m._color_raw(dirs, geo16, lidar, out_ld=3)
            t_col = timed(heads, args.iters)
            t_cmp = timed(lambda: rm.composite_rays_train(sigma, res["c"], deltas, rays, 1e-2), args.iters)
            M = xyzs.shape[0]
            out.append(dict(op="run_cuda_stages", kind=kind, fill=fill, N=N, M=M, march_ms=t_march, density_ms=t_den,
                            heads_ms=t_col, composite_ms=t_cmp,
                            heads_Gsamples_s=M / t_col / 1e6 if t_col > 0 else None,
                            density_Msamples_s=M / t_den / 1e3 if t_den > 0 else None))
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
