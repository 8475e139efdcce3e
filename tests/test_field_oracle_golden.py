"""Pins the CPU field oracle (oracle/field_oracle.py, a restatement) against outputs of the
reference's own field / renderer modules imported by path (oracle/make_golden_field.py)."""
import os

import numpy as np
import pytest
import torch

import field_cases as FC
from conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "field_ref.npz"))


@pytest.fixture(scope="module")
def orc():
    from oracle.field_oracle import FieldOracle
    return FieldOracle(FC.oracle_config(), FC.oracle_params())


def test_density_and_flow_all_time_branches(gold, orc):
    x = torch.from_numpy(gold["x"])
    for ti, t in enumerate(gold["times"].tolist()):   # first / last / interior frame, integer + fractional slice
        for lidar in (True, False):
            k = f"den_t{ti}_{'l' if lidar else 'c'}_"
            with torch.no_grad():
                r = orc.density(x, t, lidar)
            # sigma = trunc_exp(fp16 logit): compare the logits, like the other fp16 outputs of the MLPs, in
            # fp16 spacings (the golden run evaluates the flow MLP under fp16 autocast, the stand-in MLPs
            # store fp16 activations: an fp32 last-bit difference may round to the neighbouring fp16 value)
            FC.assert_fp16_close(np.log(r["sigma"].numpy()), np.log(gold[k + "sigma"]), (t, lidar, "logit"))
            FC.assert_fp16_close(r["geo_feat"].numpy(), gold[k + "geo"], (t, lidar, "geo"))
        with torch.no_grad():
            f = orc.flow(x, t)
        got = torch.cat([f["flow_forward"], f["flow_backward"]], -1).numpy()
        FC.assert_fp16_close(got, gold[f"flow_t{ti}"], (t, "flow"))
        assert np.abs(gold[f"flow_t{ti}"]).max() > 1e-3   # the flow branch is exercised


def test_color_heads_masked(gold, orc):
    geo = torch.from_numpy(gold["den_t1_l_geo"]); d = torch.from_numpy(gold["col_d"])
    mask = torch.from_numpy(gold["col_mask"])
    with torch.no_grad():
        l = orc.color(d, geo, True, mask).numpy(); c = orc.color(d, geo, False, mask).numpy()
    assert l.shape == (256, 2) and c.shape == (256, 3)
    assert np.abs(l - gold["col_l"]).max() < 1e-5 and np.abs(c - gold["col_c"]).max() < 1e-5
    assert not l[~gold["col_mask"]].any() and (l[gold["col_mask"]] > 0).all()


@pytest.mark.parametrize("ds", [1, 60])
@pytest.mark.parametrize("lidar", [True, False])
@pytest.mark.parametrize("perturb", [0, 1])
def test_run_uniform_renderer(gold, ds, lidar, perturb):
    from oracle.field_oracle import FieldOracle
    from oracle import raymarching_oracle as RO
    orc = FieldOracle(FC.oracle_config(density_scale=float(ds)), FC.oracle_params())
    k = f"ds{ds}_run_{'l' if lidar else 'c'}{perturb}_"
    o, d = gold[k + "o"], gold[k + "d"]
    nears = fars = None
    if not lidar:
        n, f = RO.near_far_from_aabb(o, d, FC.S.AABB, FC.S.MIN_NEAR)
        nears, fars = torch.from_numpy(n), torch.from_numpy(f)
    noise = torch.from_numpy(gold[k + "noise"]) if perturb else None
    with torch.no_grad():
        r = orc.run(torch.from_numpy(o), torch.from_numpy(d), 0.3, lidar, 40, nears, fars, noise)
    for name in ("depth", "image", "weights_sum", "z_vals"):
        assert FC.rel_err(r[name], gold[k + name].reshape(r[name].shape), floor=1e-4) < 1e-4, name
    assert np.abs(r["weights"].numpy() - gold[k + "weights"]).max() < 2e-5
    if ds == 60:   # saturating rays: part of the samples fall under the colour mask threshold
        frac = (gold[k + "weights"] > 1e-4).mean()
        assert 0.02 < frac < 0.98
