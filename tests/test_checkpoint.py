"""Reference checkpoint layout: the key names, shapes and optimiser groups that
`NeRFNetwork.load_reference_state_dict` / `.get_params` mirror are pinned by
tests/golden/state_dict_manifest.json, which oracle/make_golden_state_dict.py wrote from the
reference's own `NeRFNetwork.state_dict()` / `.get_params()` (network_dynamic.py:335-357,
utils.py:610-747).  CPU part: the loader consumes exactly the reference's hot-path keys and
round-trips; GPU part: a reference-layout state dict renders like the flat parameters it was cut from
and like the reference modules' own outputs (tests/golden/field_ref.npz)."""
import json
import os
import re

import numpy as np
import pytest

import field_cases as FC
from conftest import GOLDEN

torch = pytest.importorskip("torch")
S = FC.S


@pytest.fixture(scope="module")
def manifest():
    return json.load(open(os.path.join(GOLDEN, "state_dict_manifest.json")))


def _cpu_model(pkg):
    return pkg.NeRFNetwork(device="cpu", time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                           min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)


def reference_layout_state_dict(p, manifest):
    """A state dict with the reference's keys / shapes whose hot-path tensors are cut out of the flat
    oracle-layout parameters `p` exactly as oracle/ref_import.load_params_into fills the reference modules."""
    cfg = FC.oracle_config()
    sd = {}
    for k, shape, _ in manifest["keys"]:
        sd[k] = torch.zeros(*shape)     # everything the hot path does not read (un-suffixed encoders, aabb)
    sd["aabb_train"] = torch.tensor([-S.BOUND] * 3 + [S.BOUND] * 3)
    sd["aabb_infer"] = sd["aabb_train"].clone()
    from oracle.field_oracle import PLANE_COMBS
    F, Tn = cfg.n_features_hash, cfg.time_resolution
    for m in ("lidar", "camera"):
        sd[f"hash_encoder_{m}.hash_static.params"] = p[m]["hash_static"]
        base = 0
        for pi in range(3):
            per = cfg.dyn_entries[pi] * F
            for k in range(Tn):
                sd[f"hash_encoder_{m}.hash_dynamic.{pi}.hash_t.{k}.params"] = p[m]["hash_dynamic"][base:base + per]
                base += per
        off = 0
        for s, r in enumerate(cfg.plane_res):
            for ci, (a, b) in enumerate(PLANE_COMBS):
                n = cfg.n_features_plane * r[a] * r[b]
                sd[f"planes_encoder_{m}.planes.{s}.{ci}"] = p[m]["planes"][off:off + n].view(1, cfg.n_features_plane,
                                                                                         r[b], r[a])
                off += n
    sd["flow_net.grid_enc.params"] = p["flow_grid"]
    off = 0
    for li, (o, k) in zip((0, 2, 4), ((64, 32), (64, 64), (6, 64))):
        sd[f"flow_net.mlp.{li}.weight"] = p["flow_mlp"][off:off + o * k].view(o, k)
        off += o * k
    for name in ("sigma_net", "intensity_net", "raydrop_net", "color_net"):
        sd[f"{name}.params"] = p[name]
    for k, shape, _ in manifest["keys"]:
        assert tuple(sd[k].shape) == tuple(shape), k
    return sd


def test_loader_keys_are_the_reference_keys(pkg, manifest):
    m = _cpu_model(pkg)
    ref = {k: tuple(s) for k, s, _ in manifest["keys"]}
    mine = {k: tuple(shape) for k, _, _, _, shape in m._reference_keys()}
    assert all(ref.get(k) == s for k, s in mine.items()), [k for k, s in mine.items() if ref.get(k) != s][:5]
    # what the loader does not consume must be outside the optimised hot path: the reference's
    # parameter groups (network_dynamic.py:335-357) name every key it has to take
    optimised = {n for g in manifest["optimizer_groups"] for n in g["params"]}
    optimised = {n for n in optimised if ref[n] != (0,)}          # tcnn encodings without parameters
    assert optimised == set(mine), sorted(optimised ^ set(mine))[:5]
    rest = set(ref) - set(mine) - {"aabb_train", "aabb_infer"}
    assert all(re.match(r"(planes_encoder|hash_encoder)\.|view_encoder_", k) for k in rest), sorted(rest)[:5]


def test_param_groups_match_the_reference(pkg, manifest):
    m = _cpu_model(pkg)
    name_of = {id(p): n for n, p in m.named_parameters()}
    mine = {}
    for g in m.get_params(1.0):
        for p in g["params"]:
            mine[name_of[id(p)]] = g["lr"]
    key_to_param = {k: n for k, n, _, _, _ in m._reference_keys()}
    for g in manifest["optimizer_groups"]:
        for k in g["params"]:
            if k in key_to_param:
                assert mine[key_to_param[k]] == pytest.approx(g["lr"]), k
    assert pkg.optim.LR_SCALE == {n: 0.1 for n, lr in mine.items() if lr == pytest.approx(0.1)}


def test_state_dict_round_trip_cpu(pkg, manifest):
    p = FC.oracle_params()
    sd = reference_layout_state_dict(p, manifest)
    m = _cpu_model(pkg)
    missing, unexpected = m.load_reference_state_dict({"model": sd, "epoch": 3})
    assert missing == []
    assert all(re.match(r"(planes_encoder|hash_encoder)\.|view_encoder_", k) for k in unexpected)
    for mod in ("lidar", "camera"):
        for k in ("hash_static", "hash_dynamic", "planes"):
            assert torch.equal(getattr(m, f"{k}_{mod}").detach(), p[mod][k]), (mod, k)
    for k in ("flow_grid", "flow_mlp", "sigma_net", "intensity_net", "raydrop_net", "color_net"):
        assert torch.equal(getattr(m, k).detach(), p[k]), k
    out = m.reference_state_dict()
    for k, v in out.items():
        assert torch.equal(v, sd[k]), k
    del sd["sigma_net.params"]
    with pytest.raises(pkg._lib.NvsfError):
        m.load_reference_state_dict(sd)
    assert m.load_reference_state_dict(sd, strict=False)[0] == ["sigma_net.params"]


@pytest.mark.gpu
def test_reference_checkpoint_renders_like_the_reference(pkg, manifest):
    gold = np.load(os.path.join(GOLDEN, "field_ref.npz"))
    p = FC.oracle_params()
    kw = dict(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND, min_near=S.MIN_NEAR,
              min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)
    a = pkg.NeRFNetwork(**kw).eval()
    a.load_reference_state_dict({"model": reference_layout_state_dict(p, manifest)})
    b = pkg.NeRFNetwork(**kw).eval()
    b.load_flat_params(p)
    x = torch.from_numpy(gold["x"]).cuda()
    with torch.no_grad():
        for ti in (1, 2):
            t = float(gold["times"][ti])
            for lidar in (True, False):
                ra, rb = a.density(x, t, lidar), b.density(x, t, lidar)
                assert torch.equal(ra["sigma"], rb["sigma"]) and torch.equal(ra["geo_feat"], rb["geo_feat"])
                ref = gold[f"den_t{ti}_{'l' if lidar else 'c'}_sigma"]
                np.testing.assert_allclose(ra["sigma"].cpu().numpy(), ref, rtol=1e-2, atol=1e-4)
