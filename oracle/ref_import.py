"""TEST INFRASTRUCTURE (build container only — /root/reference does not exist on the GPU box).

Imports the reference's OWN field / renderer modules by file path, unchanged, with
`sys.modules` stubs for what is missing here: `tinycudann` -> oracle/tcnn_standin.py,
`trimesh` -> empty module, `nvsf.nerf.raymarching.raymarching` -> the oracle's numpy
near_far_from_aabb wrapped for torch (the only op `NeRFRenderer.run` calls,
renderer_dynamic.py:148).  Used by oracle/make_golden_field.py to generate fixtures."""
import importlib.util
import os
import sys
import types

import torch

REF = "/root/reference/nvsf/nerf"


def available():
    return os.path.isdir(REF)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference():
    from . import raymarching_oracle as RO
    from . import tcnn_standin

    sys.modules["tinycudann"] = tcnn_standin
    sys.modules.setdefault("trimesh", types.ModuleType("trimesh"))
    for pkg in ("nvsf", "nvsf.nerf", "nvsf.nerf.models", "nvsf.nerf.raymarching"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    rm = types.ModuleType("nvsf.nerf.raymarching.raymarching")

    def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
        n, f = RO.near_far_from_aabb(rays_o.detach().numpy(), rays_d.detach().numpy(), aabb.numpy(), min_near)
        return torch.from_numpy(n), torch.from_numpy(f)

    rm.near_far_from_aabb = near_far_from_aabb
    sys.modules["nvsf.nerf.raymarching.raymarching"] = rm
    sys.modules["nvsf.nerf.raymarching"].raymarching = rm
    _load("nvsf.nerf.activation", os.path.join(REF, "activation.py"))
    _load("nvsf.nerf.models.renderer_dynamic", os.path.join(REF, "models", "renderer_dynamic.py"))
    _load("nvsf.nerf.models.planes_field", os.path.join(REF, "models", "planes_field.py"))
    _load("nvsf.nerf.models.hash_field", os.path.join(REF, "models", "hash_field.py"))
    _load("nvsf.nerf.models.flow_field", os.path.join(REF, "models", "flow_field.py"))
    unet = types.ModuleType("nvsf.nerf.models.unet")  # post-hoc CNN, out of scope; not run

    class UNet(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    unet.UNet = UNet
    sys.modules["nvsf.nerf.models.unet"] = unet
    return _load("nvsf.nerf.models.network_dynamic", os.path.join(REF, "models", "network_dynamic.py"))


def load_params_into(model, cfg, p):
    """Copy flat oracle-layout parameters (oracle/field_init.py) into a reference NeRFNetwork."""
    from .field_oracle import PLANE_COMBS

    F, Tn = cfg.n_features_hash, cfg.time_resolution
    with torch.no_grad():
        for m in ("lidar", "camera"):
            he = getattr(model, f"hash_encoder_{m}")
            he.hash_static.params.copy_(p[m]["hash_static"])
            base = 0
            for pi in range(3):
                per = cfg.dyn_entries[pi] * F
                for k in range(Tn):
                    he.hash_dynamic[pi].hash_t[k].params.copy_(p[m]["hash_dynamic"][base:base + per])
                    base += per
            pe = getattr(model, f"planes_encoder_{m}")
            off = 0
            for s, r in enumerate(cfg.plane_res):
                for ci, (a, b) in enumerate(PLANE_COMBS):
                    n = cfg.n_features_plane * r[a] * r[b]
                    pe.planes[s][ci].copy_(p[m]["planes"][off:off + n].view(1, cfg.n_features_plane, r[b], r[a]))
                    off += n
        model.flow_net.grid_enc.params.copy_(p["flow_grid"])
        off = 0
        for li in (0, 2, 4):
            w = model.flow_net.mlp[li].weight
            w.copy_(p["flow_mlp"][off:off + w.numel()].view_as(w))
            off += w.numel()
        for name in ("sigma_net", "intensity_net", "raydrop_net", "color_net"):
            getattr(model, name).params.copy_(p[name])
