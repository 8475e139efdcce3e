"""TEST INFRASTRUCTURE.  Deterministic (numpy PCG64) parameters for the field in the flat
layout shared by the oracle and the product:

  per modality m in (lidar, camera):
    hash_static   tcnn flat layout of `hash_encoder_m.hash_static.params`  [level][entry][F]
    hash_dynamic  for plane p in (xy, xz, yz), for time slice k: tcnn flat layout of
                  `hash_encoder_m.hash_dynamic.p.hash_t.k.params`
    planes        for scale s, for plane c in combinations(range(4), 2): the reference tensor
                  `planes_encoder_m.planes.s.c` ([1, F, res_b, res_a], channel first) flattened
  shared:
    flow_grid     `flow_net.grid_enc.params`
    flow_mlp      `flow_net.mlp.{0,2,4}.weight` ([out, in]) concatenated
    sigma_net / intensity_net / raydrop_net / color_net   tcnn `params` (row-major [out, in_padded])

`style`: 'init' mimics the reference initialisers (tables U(-1e-4,1e-4), space planes U(0.1,0.5),
time planes 1, Xavier-uniform MLPs, last flow layer N(0,1e-3)); 'trained' uses larger tables
(U(-1,1)), perturbed time planes and a stronger flow so that every branch matters numerically.

Conditioning note.  The warped queries evaluate the dynamic hash grids at x + flow(x); at the
finest level (resolution 32768) one cell is 3e-5 wide, so ANY half-precision evaluation of the
flow MLP (the reference runs it under fp16 autocast, tcnn runs everything in fp16) moves the
query by a visible fraction of a cell.  With white-noise tables that makes the output
discontinuous in the flow and no two implementations can agree to 1e-2.  'trained' therefore
lets the amplitude of the DYNAMIC hash tables decay with the level (0.5^level), as the tables
of a trained multiresolution grid do, which keeps the comparison well conditioned without
removing any branch.
"""
import numpy as np
import torch

from .field_oracle import PLANE_COMBS


def make_params(cfg, seed=0, style="trained", as_torch=True):
    rng = np.random.default_rng(seed)
    sz = cfg.sizes()
    amp = 1e-4 if style == "init" else 1.0

    def table(n):
        return (rng.random(n, dtype=np.float32) * 2 - 1) * np.float32(amp)

    def planes():
        chunks = []
        for r in cfg.plane_res:
            for (a, b) in PLANE_COMBS:
                n = cfg.n_features_plane * r[a] * r[b]
                if 3 in (a, b):
                    v = np.ones(n, np.float32)
                    if style != "init":
                        v += (rng.random(n, dtype=np.float32) - 0.5) * np.float32(0.6)
                else:
                    v = rng.random(n, dtype=np.float32) * np.float32(0.4) + np.float32(0.1)
                    if style != "init":
                        v *= np.float32(2.0)
                chunks.append(v)
        return np.concatenate(chunks)

    def xavier(shapes, last_std=None):
        chunks = []
        for li, (o, i) in enumerate(shapes):
            if last_std is not None and li == len(shapes) - 1:
                chunks.append(rng.normal(0, last_std, size=o * i).astype(np.float32))
            else:
                b = np.sqrt(6.0 / (o + i))
                chunks.append(((rng.random(o * i, dtype=np.float32) * 2 - 1) * np.float32(b)))
        return np.concatenate(chunks)

    def dyn_table():
        v = table(sz["hash_dynamic"])
        if style == "init":
            return v
        F, Tn, off = cfg.n_features_hash, cfg.time_resolution, 0
        for pi in range(3):
            for _ in range(Tn):
                for l, lv in enumerate(cfg.dyn_levels[pi]):
                    a, b = off + lv["offset"] * F, off + (lv["offset"] + lv["size"]) * F
                    v[a:b] *= np.float32(0.5 ** l)
                off += cfg.dyn_entries[pi] * F
        return v

    h = cfg.hidden
    fin = cfg.flow_levels * cfg.flow_features // 4
    p = {}
    for m in ("lidar", "camera"):
        p[m] = dict(hash_static=table(sz["hash_static"]), hash_dynamic=dyn_table(), planes=planes())
    p["flow_grid"] = table(sz["flow_grid"])
    p["flow_mlp"] = xavier([(h, fin), (h, h), (6, h)], last_std=1e-3)
    p["sigma_net"] = xavier([(h, 128), (16, h)])
    if style != "init":
        p["sigma_net"] *= np.float32(0.5)
    p["intensity_net"] = xavier([(h, 96), (h, h), (16, h)])
    p["raydrop_net"] = xavier([(h, 96), (h, h), (16, h)])
    p["color_net"] = xavier([(h, 32), (h, h), (16, h)])
    for k, v in list(p.items()):
        if isinstance(v, dict):
            for kk, vv in v.items():
                assert vv.size == sz[kk], (kk, vv.size, sz[kk])
        else:
            assert v.size == sz[k], (k, v.size, sz[k])
    if as_torch:
        p = {k: ({kk: torch.from_numpy(vv) for kk, vv in v.items()} if isinstance(v, dict) else torch.from_numpy(v))
             for k, v in p.items()}
    return p
