"""TEST INFRASTRUCTURE (build container only).  Generates tests/golden/field_ref.npz by running
the reference's OWN field / renderer code (imported by file path, oracle/ref_import.py, with the
tinycudann stand-in) on seeded inputs and on parameters from oracle/field_init.py (seed 0,
style 'trained').  Usage:  python -m oracle.make_golden_field"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import field_init, ref_import  # noqa: E402
from oracle import raymarching_oracle as RO  # noqa: E402
from oracle.field_oracle import FieldConfig  # noqa: E402

S = importlib.import_module("selfsupervised-nvsf_b200.synth")
TIMES = [0.0, 0.5, 31.0 / 63.0, 1.0, 0.2]


def config(density_scale=1.0):
    return FieldConfig(bound=S.BOUND, num_frames=S.NUM_FRAMES, time_resolution=S.TIME_RESOLUTION,
                       min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH,
                       density_scale=density_scale)


def main():
    nd = ref_import.import_reference()
    out = {}
    cfg = config()
    p = field_init.make_params(cfg, seed=0, style="trained")
    rng = np.random.default_rng(123)
    x = ((rng.random((256, 3), dtype=np.float32) * 2 - 1) * np.float32(1.95)).astype(np.float32)
    x[:8] = np.float32(S.BOUND) * np.sign(x[:8])          # points clipped onto the box faces
    out["x"] = x
    for ds in (1.0, 60.0):
        model = nd.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                               min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR,
                               lidar_max_depth=S.LIDAR_MAX_DEPTH, density_scale=ds).eval()
        ref_import.load_params_into(model, cfg, p)
        tag = f"ds{int(ds)}_"
        # the reference runs its model under fp16 autocast (configs/kitti360_1908.txt:23, trainer.py:929,1139,1318)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.float16):
            if ds == 1.0:
                for ti, t in enumerate(TIMES):
                    tt = torch.tensor([[t]], dtype=torch.float32)
                    for lidar in (True, False):
                        r = model.density(torch.from_numpy(x), tt, lidar)
                        k = f"den_t{ti}_{'l' if lidar else 'c'}_"
                        out[k + "sigma"] = r["sigma"].numpy(); out[k + "geo"] = r["geo_feat"].numpy()
                    f = model.flow(torch.from_numpy(x), tt)
                    out[f"flow_t{ti}"] = torch.cat([f["flow_forward"], f["flow_backward"]], -1).numpy()
                out["times"] = np.array(TIMES, np.float32)
                # color heads on given geo features / directions, with a mask
                geo = torch.from_numpy(out["den_t1_l_geo"]); d = torch.from_numpy(S.lidar_rays(256, seed=9)[1])
                mask = torch.from_numpy(rng.random(256) < 0.6)
                model.out_dim = 2
                out["col_l"] = model.color(torch.from_numpy(x), d, True, mask, geo).numpy()
                model.out_dim = 3
                out["col_c"] = model.color(torch.from_numpy(x), d, False, mask, geo).numpy()
                out["col_d"] = d.numpy(); out["col_mask"] = mask.numpy()
            for lidar in (True, False):
                o, d = (S.lidar_rays if lidar else S.camera_rays)(48, seed=5)
                for perturb in (False, True):
                    torch.manual_seed(77)
                    noise = torch.rand(48, 40).numpy() if perturb else None
                    torch.manual_seed(77)
                    r = model.render(torch.from_numpy(o)[None], torch.from_numpy(d)[None], torch.tensor([[0.3]]),
                                     cal_lidar_color=lidar, staged=False, num_steps=40, perturb=perturb)
                    sfx = "_lidar" if lidar else ""
                    k = tag + f"run_{'l' if lidar else 'c'}{int(perturb)}_"
                    out[k + "o"], out[k + "d"] = o, d
                    if noise is not None:
                        out[k + "noise"] = noise
                    out[k + "depth"] = r["depth" + sfx].numpy().reshape(-1)
                    out[k + "image"] = r["image" + sfx].numpy().reshape(48, -1)
                    out[k + "weights_sum"] = r["weights_sum" + sfx].numpy()
                    out[k + "weights"] = r["weights"].numpy(); out[k + "z_vals"] = r["z_vals"].numpy()
    path = os.path.join(ROOT, "tests", "golden", "field_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
