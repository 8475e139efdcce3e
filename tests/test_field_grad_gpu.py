"""Parity of the sm_100a training path (forward with kept intermediates + backward to the fp32
master parameters) against the reference's own autograd gradients (tests/golden/field_grad_ref.npz,
oracle/make_golden_grad.py) and against the CPU oracle.

Tolerance: the backward runs its GEMM operands in fp16 (device-chosen power-of-two scale) like the forward (BASELINE.json
north_star: 1e-2 for MLP outputs); gradient tensors are compared in relative L2 norm over the
sampled reference entries, over the whole tensor norm and over its sum."""
import os

import numpy as np
import pytest

import field_cases as FC
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
S = FC.S

# Cases of 256 rays x 64 samples (oracle/make_golden_grad.py).  Tolerance per tensor: 1e-2 (flow MLP 2e-2; flow grid
# 5e-2: each of its entries is touched by about one sample, so a ReLU of the flow MLP that flips between two
# evaluations is not averaged out, see tests/test_loss_terms_gpu.py::test_flow_loss_matches_oracle), or
# TWICE the conditioning floor of the reference gradient where that is larger — FC.oracle_grad_floor measures, in
# the test itself, how far the CPU oracle's own gradient moves when the ray origins change by one fp32 ulp
# (measured 1-2.6 % on these cases: finest hash level = 32768 cells per unit, fp16 activations).
RTOL = {"hash_static": 1e-2, "hash_dynamic": 1e-2, "planes": 1e-2, "flow_grid": 5e-2, "flow_mlp": 2e-2,
        "sigma_net": 1e-2, "intensity_net": 1e-2, "raydrop_net": 1e-2, "color_net": 1e-2}
FLOOR_FACTOR = 2.0


def tolerance(case, tag, name):
    return max(RTOL[name], FLOOR_FACTOR * FC.oracle_grad_floor(case, tag)[name])


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "field_grad_ref.npz"))


def make_model(pkg, ds):
    m = pkg.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                        min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH,
                        density_scale=ds)
    m.load_flat_params(FC.oracle_params())
    return m.train()


def run_case(m, case):
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    sfx = "_lidar" if case["lidar"] else ""
    n_steps = case["coef"]["c"].shape[1]
    out = m.render(dev(case["o"])[None], dev(case["d"])[None], torch.tensor([[case["t"]]], device="cuda"),
                   cal_lidar_color=case["lidar"], staged=False, num_steps=n_steps,
                   noise=None if case["noise"] is None else dev(case["noise"]))
    std = dict(depth=out["depth" + sfx], image=out["image" + sfx], weights=out["weights"],
               weights_sum=out["weights_sum" + sfx])
    loss = FC.linear_loss(std, case["coef"], to=dev)
    return loss, std


def grads_of(m, lidar):
    mod = "lidar" if lidar else "camera"
    g = {}
    for name in FC.GRAD_NAMES:
        p = getattr(m, f"{name}_{mod}" if name in ("hash_static", "hash_dynamic", "planes") else name)
        g[name] = np.zeros(p.numel(), np.float32) if p.grad is None else p.grad.detach().cpu().numpy().reshape(-1)
    return g


@pytest.mark.parametrize("tag", FC.GRAD_CASES)
def test_parameter_gradients_match_reference(pkg, gold, tag):
    case = FC.grad_case(gold, tag)
    m = make_model(pkg, case["ds"])
    loss, out = run_case(m, case)
    assert abs(loss.item() - float(gold[tag + "_loss"])) < 1e-2 * max(1.0, abs(float(gold[tag + "_loss"])))
    loss.backward()
    torch.cuda.synchronize()
    g = grads_of(m, case["lidar"])
    for name in FC.GRAD_NAMES:
        assert np.isfinite(g[name]).all(), name
        FC.check_grad_summary(gold, tag, name, g[name], tolerance(case, tag, name), "cuda")
    # the other modality's encoders are untouched
    other = "camera" if case["lidar"] else "lidar"
    assert getattr(m, f"hash_static_{other}").grad is None


def test_gradients_match_cpu_oracle_elementwise(pkg, gold):
    """Full-tensor comparison against the oracle's autograd gradients (the fixture only keeps a
    sample of the reference's): relative L2 error per parameter tensor."""
    case = FC.grad_case(gold, "l_mid")
    m = make_model(pkg, case["ds"])
    loss, _ = run_case(m, case)
    loss.backward()
    g = grads_of(m, True)
    e, _, _ = FC.oracle_grads(case)
    for name in FC.GRAD_NAMES:
        ref = e[name].reshape(-1).astype(np.float64)
        if not ref.any():
            assert not g[name].any(), name
            continue
        err = np.linalg.norm(g[name] - ref) / np.linalg.norm(ref)
        assert err < tolerance(case, "l_mid", name), (name, err)
        # Sparsity: the stand-in's table gradients pass through an fp16 cast (like tcnn's half
        # atomics), which zeroes entries below 6e-8; this backward keeps them in fp32.  Nothing
        # larger than that may appear where the reference has an exact zero.
        if name in ("hash_static", "hash_dynamic", "flow_grid"):
            extra = g[name][ref == 0]
            assert np.linalg.norm(extra) < 1e-3 * np.linalg.norm(ref) and np.abs(extra).max() < 1e-6, name
            # ... and what the reference has where this backward has an exact zero (rows whose scaled fp16 output
            # gradient is zero are skipped) is negligible
            assert np.linalg.norm(ref[g[name] == 0]) < 1e-3 * np.linalg.norm(ref), name


def test_backward_is_linear_in_the_output_gradient(pkg, gold):
    """Size-independent property: backward(a*g1 + b*g2) == a*backward(g1) + b*backward(g2)."""
    case = FC.grad_case(gold, "c_mid")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def grads(cd, ci):
        m = make_model(pkg, case["ds"])
        out = m.render(dev(case["o"])[None], dev(case["d"])[None], torch.tensor([[case["t"]]], device="cuda"),
                       cal_lidar_color=False, staged=False, num_steps=32)
        (cd * out["depth"].sum() + ci * out["image"].sum()).backward()
        return grads_of(m, False)

    g1, g2, g3 = grads(1.0, 0.0), grads(0.0, 1.0), grads(2.0, -3.0)
    for name in ("hash_static", "planes", "sigma_net", "color_net", "flow_mlp"):
        want = 2.0 * g1[name].astype(np.float64) - 3.0 * g2[name]
        err = np.linalg.norm(g3[name] - want) / max(np.linalg.norm(want), 1e-30)
        assert err < 2e-2, (name, err)


def test_fused_grad_accumulation_and_repeat(pkg, gold):
    """Accumulating straight into existing .grad buffers gives the same result as autograd's
    accumulation, and two backward passes accumulate."""
    case = FC.grad_case(gold, "l_first")
    m1 = make_model(pkg, case["ds"])
    run_case(m1, case)[0].backward()
    g1 = grads_of(m1, True)
    m2 = make_model(pkg, case["ds"])
    m2.fused_grad_accumulation = True
    for p in m2.parameters():
        p.grad = torch.zeros_like(p)
    run_case(m2, case)[0].backward()
    run_case(m2, case)[0].backward()
    g2 = grads_of(m2, True)
    for name in FC.GRAD_NAMES:
        a, b = 2.0 * g1[name].astype(np.float64), g2[name].astype(np.float64)
        assert np.linalg.norm(a - b) <= 1e-3 * max(np.linalg.norm(a), 1e-30), name


def test_train_forward_equals_inference_forward(pkg, gold):
    """The training forward runs the same staged pipeline as inference but keeps its intermediates
    (un-fused gather stage writing the feature rows): bit-identical in density_mode 1; in the default
    mode 2 inference accumulates the sigma net inside the fused tcgen05 kernel, so the two agree to
    fp32 accumulation-order rounding."""
    L = pkg._lib.lib()
    case = FC.grad_case(gold, "l_mid")
    prev = L.nvsf_density_mode_get()
    try:
        for mode in (1, 2):
            assert L.nvsf_set_option(b"density_mode", mode) == 0
            m = make_model(pkg, case["ds"])
            _, out = run_case(m, case)
            with torch.no_grad():
                _, ref = run_case(m, case)
            for k in ("depth", "image", "weights", "weights_sum"):
                if mode == 1:
                    assert torch.equal(out[k].detach(), ref[k]), k
                else:
                    torch.testing.assert_close(out[k].detach(), ref[k], rtol=1e-3, atol=1e-5, msg=k)
            assert out["depth"].requires_grad and not ref["depth"].requires_grad
    finally:
        L.nvsf_set_option(b"density_mode", prev)


@pytest.mark.parametrize("tag", ["l_mid", "c_mid"])
def test_mlp_backward_tcgen05_matches_mma_sync(pkg, gold, tag):
    """The MLP backward on tcgen05 (csrc/mlp_bwd_tc.cuh: activations as the A operand from tensor memory, weight
    gradients accumulated in tensor memory from MN-major views of the sample tiles; option mlp_bwd_tc, default)
    against the mma.sync kernel k_mlp_bwd: the same fp16 operands and scales, fp32 sums in a different order."""
    L = pkg._lib.lib()
    case = FC.grad_case(gold, tag)
    g = {}
    try:
        for tc in (0, 1):
            assert L.nvsf_set_option(b"mlp_bwd_tc", tc) == 0 and L.nvsf_get_option(b"mlp_bwd_tc") == tc
            m = make_model(pkg, case["ds"])
            run_case(m, case)[0].backward()
            g[tc] = grads_of(m, case["lidar"])
    finally:
        L.nvsf_set_option(b"mlp_bwd_tc", 1)
    errs = {}
    for name in FC.GRAD_NAMES:
        if name not in g[0]:
            continue
        a, b = g[1][name].astype(np.float64), g[0][name].astype(np.float64)
        errs[name] = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    print(tag, {k: f"{v:.2e}" for k, v in errs.items()})
    for name, e in errs.items():
        assert e < 5e-3, (name, e, errs)


@pytest.mark.parametrize("lidar,N,Sn", [(True, 5, 7), (False, 3, 50), (True, 2, 300), (False, 130, 1), (True, 300, 1)])
def test_mlp_backward_tcgen05_tiny_and_ragged(pkg, lidar, N, Sn):
    """Row counts below one 128-row tile, tiles that hold several rays (the per-ray direction encodings of the LiDAR
    heads), rays longer than a tile, one sample per ray: tcgen05 backward == mma.sync backward on every gradient."""
    L = pkg._lib.lib()
    o, d = (S.lidar_rays if lidar else S.camera_rays)(N, seed=11)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    sfx = "_lidar" if lidar else ""
    g = {}
    try:
        for tc in (0, 1):
            assert L.nvsf_set_option(b"mlp_bwd_tc", tc) == 0
            m = make_model(pkg, 1.0)
            out = m.render(dev(o)[None], dev(d)[None], torch.tensor([[0.6]], device="cuda"), cal_lidar_color=lidar,
                           staged=False, num_steps=Sn)
            (out["depth" + sfx].sum() + 3.0 * out["image" + sfx].sum() + 0.5 * out["weights_sum" + sfx].sum()).backward()
            g[tc] = grads_of(m, lidar)
    finally:
        L.nvsf_set_option(b"mlp_bwd_tc", 1)
    for name in FC.GRAD_NAMES:
        a, b = g[1][name].astype(np.float64), g[0][name].astype(np.float64)
        assert np.isfinite(a).all()
        assert np.linalg.norm(a - b) <= 5e-3 * max(np.linalg.norm(b), 1e-30), (name, np.linalg.norm(a - b), np.linalg.norm(b))


# Small cases do not average out the fp16 ReLU flips between two correct evaluations: measured against the CPU oracle on a
# B200 (tools/dbg_ragged.py; identical for the tcgen05 and the mma.sync backward), aligned and ragged shapes alike:
#   rays x samples   hash_static  hash_dynamic  planes  flow_grid  flow_mlp  sigma_net
#   32 x 32          0.013        0.016         0.005   0.021      0.016     0.003
#   37 x 37          0.012        0.013         0.006   0.040      0.038     0.006
#   150 x 37         0.018        0.020         0.008   0.037      0.019     0.005
#   128 x 128        0.015        0.014         0.003   0.024      0.005     0.003
#   21 x 150 (cam)   0.014        0.014         0.004   0.019      0.013     0.003
# A row-indexing mistake on a ragged tile gives errors of order one; the tolerances below sit between the two, and the
# additivity test after this one pins ragged against aligned tilings of the SAME kernels at 5e-3.
RAGGED_RTOL = {"flow_grid": 8e-2, "flow_mlp": 6e-2}


def _ragged_case(lidar, N, Sn, seed=5):
    rng = np.random.default_rng(23)
    o, d = (S.lidar_rays if lidar else S.camera_rays)(N, seed=seed)
    nch = 2 if lidar else 3
    return dict(lidar=lidar, t=0.45, ds=1.0, o=o, d=d, noise=None,
                coef=dict(a=rng.normal(size=N).astype(np.float32), b=rng.normal(size=(N, nch)).astype(np.float32),
                          c=(0.1 * rng.normal(size=(N, Sn))).astype(np.float32), e=rng.normal(size=N).astype(np.float32)))


@pytest.mark.parametrize("lidar,N,Sn", [(True, 37, 37), (False, 21, 150)])
def test_ragged_shapes_match_cpu_oracle(pkg, lidar, N, Sn):
    """Row counts that are no multiple of a warp or of a 128-row tile, tiles that straddle several rays (37 x 37), rays
    that straddle tiles (21 x 150): every parameter gradient of the default (tcgen05) backward against the CPU oracle's
    autograd."""
    case = _ragged_case(lidar, N, Sn)
    m = make_model(pkg, case["ds"])
    loss, _ = run_case(m, case)
    loss.backward()
    g = grads_of(m, lidar)
    e, eloss, _ = FC.oracle_grads(case)
    assert abs(float(loss.detach()) - eloss) <= 1e-3 * max(abs(eloss), 1.0), (float(loss.detach()), eloss)
    for name in FC.GRAD_NAMES:
        ref = e[name].reshape(-1).astype(np.float64)
        if not ref.any():
            assert not g[name].any(), name
            continue
        err = np.linalg.norm(g[name] - ref) / np.linalg.norm(ref)
        assert err < RAGGED_RTOL.get(name, 3e-2), (name, err)


@pytest.mark.parametrize("lidar,Sn,Na,Nb", [(True, 37, 37, 91), (False, 150, 21, 43), (True, 1, 100, 156)])
def test_gradients_are_additive_over_ragged_and_aligned_batches(pkg, lidar, Sn, Na, Nb):
    """Size-independent property that isolates the tiling: the rays of a batch whose row count fills whole 128-row tiles
    (Na + Nb rays) are also rendered as two ragged batches (Na and Nb rays: partial last tiles, partial warps); the
    gradients of the parts must add up to the gradient of the whole.  Same kernels, same arithmetic per row — only the
    tile partition and the per-launch power-of-two fp16 scale differ."""
    assert ((Na + Nb) * Sn) % 128 == 0 and (Na * Sn) % 32 != 0
    whole = _ragged_case(lidar, Na + Nb, Sn)

    def part(lo, hi):
        c = dict(whole, o=whole["o"][lo:hi], d=whole["d"][lo:hi])
        c["coef"] = {k: v[lo:hi] for k, v in whole["coef"].items()}
        return c

    def grads(case):
        m = make_model(pkg, 1.0)
        loss, _ = run_case(m, case)
        loss.backward()
        return grads_of(m, lidar)

    gw, ga, gb = grads(whole), grads(part(0, Na)), grads(part(Na, Na + Nb))
    for name in FC.GRAD_NAMES:
        want = gw[name].astype(np.float64)
        if not want.any():
            assert not ga[name].any() and not gb[name].any(), name
            continue
        err = np.linalg.norm(ga[name].astype(np.float64) + gb[name] - want) / np.linalg.norm(want)
        assert err < 5e-3, (name, err)
