"""Pins the CPU oracle (oracle/raymarching_oracle.c) against vectors produced by the
reference's own extension (unmodified raymarching.cu built for sm_100a and run on a B200 by
oracle/make_golden_raymarching.py).  Runs without a GPU."""
import os

import numpy as np
import pytest

import cases
from conftest import GOLDEN, assert_bits_equal

S = cases.S
C, H, BOUND = S.CASCADE, S.GRID_SIZE, S.BOUND


@pytest.fixture(scope="module")
def gold():
    path = os.path.join(GOLDEN, "raymarching_ref_sm100a.npz")
    assert os.path.exists(path), "golden vectors missing"
    return np.load(path)


def close(a, b, rtol=1e-5, atol=1e-7):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def test_near_far(gold, oracle):
    n, f = oracle.near_far_from_aabb(gold["nf_o"], gold["nf_d"], S.AABB, S.MIN_NEAR)
    assert_bits_equal(n, gold["nf_nears"]); assert_bits_equal(f, gold["nf_fars"])
    assert (gold["nf_nears"] > 1e38).any() and (gold["nf_nears"] < 1e38).any()  # both hit and miss rays


def test_sph_from_ray(gold, oracle):
    close(oracle.sph_from_ray(gold["sph_o"], gold["sph_d"], 3.0), gold["sph_coords"], rtol=0, atol=2e-6)


def test_morton(gold, oracle):
    assert_bits_equal(oracle.morton3D(gold["mt_coords"]), gold["mt_indices"])
    assert_bits_equal(oracle.morton3D_invert(gold["mt_indices"]), gold["mt_back"])
    assert np.array_equal(gold["mt_back"][:600], gold["mt_coords"][:600])


def test_packbits(gold, oracle):
    assert_bits_equal(oracle.packbits(gold["pb_grid"], 0.25), gold["pb_bits"])
    assert_bits_equal(S.packbits_np(gold["pb_grid"], 0.25), gold["pb_bits"])


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_march_and_composite_train(gold, oracle, ci):
    p = f"mt{ci}_"
    _, N, max_steps, perturb, m = gold[p + "meta"].tolist()
    fill = str(gold[p + "fill"]); dt_gamma = float(gold[p + "dt_gamma"])
    bf = cases.bitfield(fill, seed=6)
    # the committed inputs are the ones the reference saw; they must also be what cases.py makes
    o, d, nears, fars, noises = cases.march_inputs(str(gold[p + "kind"]), N, seed=50 + ci, perturb=bool(perturb))
    assert_bits_equal(o, gold[p + "o"]); assert_bits_equal(d, gold[p + "d"]); assert_bits_equal(noises, gold[p + "noises"])
    assert_bits_equal(nears, gold[p + "nears"]); assert_bits_equal(fars, gold[p + "fars"])
    x, dd, l, r, c = oracle.march_rays_train(o, d, BOUND, bf, C, H, nears, fars, noises, dt_gamma=dt_gamma,
                                             max_steps=max_steps)
    assert c.tolist() == [m, N]
    assert_bits_equal(r, gold[p + "rays"], "rays"); assert_bits_equal(x[:m], gold[p + "xyzs"], "xyzs")
    assert_bits_equal(l[:m], gold[p + "deltas"], "deltas")
    assert m > 0 and r[:, 2].max() <= max_steps
    sig, rgb = cases.field_values(m, seed=int(gold[p + "sig_seed"])); sig *= np.float32(30.0)
    ws, depth, image = oracle.composite_rays_train_forward(sig, rgb, l[:m], r, 1e-4)
    close(ws, gold[p + "ws"]); close(depth, gold[p + "depth"]); close(image, gold[p + "image"])
    gs, gr = oracle.composite_rays_train_backward(gold[p + "g_ws"], gold[p + "g_im"], sig, rgb, l[:m], r,
                                                  gold[p + "ws"], gold[p + "image"], 1e-4)
    close(gs, gold[p + "grad_sigmas"], rtol=2e-5, atol=2e-5); close(gr, gold[p + "grad_rgbs"], atol=5e-6)


def test_inference_step(gold, oracle):
    n_alive, n_step = gold["inf_meta"].tolist()
    bf = cases.bitfield("shell", seed=6)
    x, d, l = oracle.march_rays(n_alive, n_step, gold["inf_alive"], gold["inf_rays_t"], gold["inf_o"], gold["inf_d"],
                                BOUND, bf, C, H, gold["inf_nears"], gold["inf_fars"], gold["inf_noises"],
                                dt_gamma=S.DT_GAMMA)
    assert_bits_equal(x, gold["inf_xyzs"]); assert_bits_equal(d, gold["inf_dirs"]); assert_bits_equal(l, gold["inf_deltas"])
    sig, rgb = cases.field_values(n_alive * n_step, seed=71); sig *= np.float32(60.0)
    a, t, ws, de, im = oracle.composite_rays(n_alive, n_step, gold["inf_alive"], gold["inf_rays_t"], sig, rgb, l,
                                             gold["inf_ws0"], gold["inf_de0"], gold["inf_im0"], 1e-2)
    assert np.array_equal(a, gold["inf_alive_out"]) and (a < 0).any() and (a >= 0).any()
    close(t, gold["inf_t_out"]); close(ws, gold["inf_ws"]); close(de, gold["inf_depth"]); close(im, gold["inf_image"])


def test_oracle_edge_cases(oracle):
    """empty inputs, empty grid, ragged counts, counter accumulation, buffer overflow drop."""
    o, d, nears, fars, noises = cases.march_inputs("lidar", 64, seed=1, perturb=True)
    empty = cases.bitfield("empty")
    x, dd, l, r, c = oracle.march_rays_train(o, d, BOUND, empty, C, H, nears, fars, noises, dt_gamma=S.DT_GAMMA)
    assert c.tolist() == [0, 64] and not r[:, 1:].any() and np.array_equal(r[:, 0], np.arange(64))
    x, dd, l, r, c = oracle.march_rays_train(o[:0], d[:0], BOUND, empty, C, H, nears[:0], fars[:0], noises[:0])
    assert c.tolist() == [0, 0]
    shell = cases.bitfield("shell")
    x, dd, l, r, c = oracle.march_rays_train(o, d, BOUND, shell, C, H, nears, fars, noises, dt_gamma=S.DT_GAMMA)
    m = int(c[0])
    assert np.array_equal(r[:, 1], np.concatenate([[0], np.cumsum(r[:, 2])[:-1]]))
    assert len(set(r[:, 2].tolist())) > 3  # ragged
    # accumulated counter
    x2, _, _, r2, c2 = oracle.march_rays_train(o, d, BOUND, shell, C, H, nears, fars, noises, dt_gamma=S.DT_GAMMA,
                                               counter=np.array([m, 0], np.int32), M=2 * m)
    assert c2.tolist() == [2 * m, 64] and np.array_equal(r2[:, 1], r[:, 1] + m)
    assert_bits_equal(x2[m:2 * m], x[:m])
    # too-small buffer: trailing rays dropped, leading ones intact
    xs, _, ls, rs, cs = oracle.march_rays_train(o, d, BOUND, shell, C, H, nears, fars, noises, dt_gamma=S.DT_GAMMA, M=m // 2)
    kept = rs[:, 1] + rs[:, 2] <= m // 2
    last = int((rs[kept][:, 1] + rs[kept][:, 2]).max())
    assert_bits_equal(xs[:last], x[:last]); assert not xs[last:].any()
    # delta invariants: deltas[:,0] is the step, deltas[:,1] the distance from the previous sample end
    assert (l[:m, 0] > 0).all() and (l[:m, 1] >= l[:m, 0] - 1e-6).all()
    # compositing: T_thresh=1 terminates after the first sample
    sig = np.full(m, 5.0, np.float32); rgb = np.ones((m, 3), np.float32)
    ws, depth, image = oracle.composite_rays_train_forward(sig, rgb, l[:m], r, 1.0)
    first = r[:, 1][r[:, 2] > 0]
    a0 = 1 - np.exp(-5.0 * l[first, 0].astype(np.float64))
    close(ws[r[:, 2] > 0], a0, rtol=1e-5); assert not ws[r[:, 2] == 0].any()
