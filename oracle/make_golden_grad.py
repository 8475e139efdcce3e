"""TEST INFRASTRUCTURE (build container only).  Generates tests/golden/field_grad_ref.npz: parameter
gradients of the reference's OWN field / renderer code (imported by file path, oracle/ref_import.py,
with the tinycudann stand-in) for a fixed linear functional of the render outputs

    L = sum(a * depth) + sum(b * image) + sum(c * weights) + sum(e * weights_sum)

on seeded rays and parameters (oracle/field_init.py, seed 0, style 'trained').  The gradient
tensors have 94 M entries, so the fixture keeps, per parameter tensor, its sum, L2 norm, number of
non-zeros and the values at <= 2048 indices (largest magnitudes + a seeded random subset of the
non-zeros).  Usage:  python -m oracle.make_golden_grad"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import field_init, ref_import  # noqa: E402
from oracle.field_oracle import PLANE_COMBS  # noqa: E402
from oracle.make_golden_field import config  # noqa: E402

S = importlib.import_module("selfsupervised-nvsf_b200.synth")
N_RAYS, N_STEPS = 256, 64   # 16 K samples per case: one fp16 ReLU sign flip no longer moves a whole tensor
CASES = [  # (tag, lidar, time, density_scale, perturb)
    ("l_mid", True, 0.3, 60.0, True),
    ("c_mid", False, 0.3, 60.0, False),
    ("l_first", True, 0.0, 1.0, False),
    ("c_last", False, 1.0, 60.0, True),
]
LOSS_SCALE = 4096.0
SHARED = ("flow_grid", "flow_mlp", "sigma_net", "intensity_net", "raydrop_net", "color_net")


def loss_coeffs(seed, n_ch):
    rng = np.random.default_rng(seed)
    return dict(a=rng.normal(size=N_RAYS).astype(np.float32),
                b=rng.normal(size=(N_RAYS, n_ch)).astype(np.float32),
                c=(rng.normal(size=(N_RAYS, N_STEPS)) * 0.3).astype(np.float32),
                e=(rng.normal(size=N_RAYS) * 0.5).astype(np.float32))


def flat_grads(model, cfg, mod):
    """Reference .grad tensors -> the flat layout of oracle/field_init.py (zeros where None)."""
    def g(p):
        return (p.grad / LOSS_SCALE if p.grad is not None else torch.zeros_like(p)).detach().reshape(-1)

    he = getattr(model, f"hash_encoder_{mod}")
    out = {"hash_static": g(he.hash_static.params)}
    out["hash_dynamic"] = torch.cat([g(he.hash_dynamic[pi].hash_t[k].params)
                                     for pi in range(3) for k in range(cfg.time_resolution)])
    pe = getattr(model, f"planes_encoder_{mod}")
    out["planes"] = torch.cat([g(pe.planes[s][ci]) for s in range(len(cfg.plane_res))
                               for ci in range(len(PLANE_COMBS))])
    out["flow_grid"] = g(model.flow_net.grid_enc.params)
    out["flow_mlp"] = torch.cat([g(model.flow_net.mlp[li].weight) for li in (0, 2, 4)])
    for name in ("sigma_net", "intensity_net", "raydrop_net", "color_net"):
        out[name] = g(getattr(model, name).params)
    return {k: v.numpy() for k, v in out.items()}


def summarise(g, seed):
    nz = np.flatnonzero(g)
    order = nz[np.argsort(-np.abs(g[nz]))]
    top = order[:512]
    rest = order[512:]
    rng = np.random.default_rng(seed)
    pick = rng.choice(rest, size=min(1536, rest.size), replace=False) if rest.size else rest
    idx = np.sort(np.concatenate([top, pick])).astype(np.int64)
    return dict(sum=np.float64(g.sum(dtype=np.float64)), l2=np.float64(np.sqrt((g.astype(np.float64) ** 2).sum())),
                nnz=np.int64(nz.size), idx=idx, val=g[idx].astype(np.float32))


def main():
    nd = ref_import.import_reference()
    out = {}
    p = field_init.make_params(config(), seed=0, style="trained")
    for ci, (tag, lidar, t, ds, perturb) in enumerate(CASES):
        cfg = config(ds)
        model = nd.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                               min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR,
                               lidar_max_depth=S.LIDAR_MAX_DEPTH, density_scale=ds).train()
        ref_import.load_params_into(model, cfg, p)
        o, d = (S.lidar_rays if lidar else S.camera_rays)(N_RAYS, seed=11 + ci)
        torch.manual_seed(5 + ci)
        noise = torch.rand(N_RAYS, N_STEPS).numpy() if perturb else None
        torch.manual_seed(5 + ci)   # NeRFRenderer.run draws the same torch.rand(N, S) when perturb
        # the reference trains under fp16 autocast with a GradScaler (configs/kitti360_1908.txt:23,
        # trainer.py:119,1318,1332): same here, with a fixed power-of-two loss scale (exact to undo)
        with torch.autocast("cpu", dtype=torch.float16):
            r = model.render(torch.from_numpy(o)[None], torch.from_numpy(d)[None], torch.tensor([[t]]),
                             cal_lidar_color=lidar, staged=False, num_steps=N_STEPS, perturb=perturb)
        sfx = "_lidar" if lidar else ""
        co = loss_coeffs(100 + ci, 2 if lidar else 3)
        loss = ((torch.from_numpy(co["a"]) * r["depth" + sfx].reshape(-1)).sum()
                + (torch.from_numpy(co["b"]) * r["image" + sfx].reshape(N_RAYS, -1)).sum()
                + (torch.from_numpy(co["c"]) * r["weights"]).sum()
                + (torch.from_numpy(co["e"]) * r["weights_sum" + sfx]).sum())
        (loss * LOSS_SCALE).backward()
        k = f"{tag}_"
        out[k + "o"], out[k + "d"] = o, d
        out[k + "meta"] = np.array([float(lidar), t, ds, float(perturb)], np.float64)
        if noise is not None:
            out[k + "noise"] = noise
        for name, v in co.items():
            out[k + "coef_" + name] = v
        out[k + "loss"] = np.float64(loss.item())
        out[k + "depth"] = r["depth" + sfx].detach().numpy().reshape(-1)
        g = flat_grads(model, cfg, "lidar" if lidar else "camera")
        for name, arr in g.items():
            sm = summarise(arr, seed=ci)
            for kk, vv in sm.items():
                out[f"{k}g_{name}_{kk}"] = vv
            print(tag, name, "nnz", int(sm["nnz"]), "l2 %.4g" % sm["l2"])
        # the other modality's encoders must receive no gradient
        other = flat_grads(model, cfg, "camera" if lidar else "lidar")
        assert not other["hash_static"].any() and not other["planes"].any()
    path = os.path.join(ROOT, "tests", "golden", "field_grad_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
