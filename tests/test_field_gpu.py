"""Parity of the sm_100a field + uniform renderer against the CPU oracle (oracle/field_oracle.py,
itself pinned to the reference modules by tests/test_field_oracle_golden.py).

Tolerances (BASELINE.json north_star): 1e-2 relative for anything that passes through the fp16
MLPs (sigma, geo features, colours, composited outputs of the uniform renderer); encoder features
are compared at fp16 resolution (they are stored as fp16 for the tensor cores)."""
import os

import numpy as np
import pytest

import field_cases as FC
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

S = FC.S
TIMES = [0.0, 0.5, 31.0 / 63.0, 1.0, 0.2]


@pytest.fixture(scope="module")
def model(pkg):
    m = pkg.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                        min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)
    m.load_flat_params(FC.oracle_params())
    return m.eval()


@pytest.fixture(scope="module")
def orc():
    from oracle.field_oracle import FieldOracle
    return FieldOracle(FC.oracle_config(), FC.oracle_params())


def pts(n, seed):
    rng = np.random.default_rng(seed)
    x = ((rng.random((n, 3), dtype=np.float32) * 2 - 1) * np.float32(1.95)).astype(np.float32)
    x[:8] = np.float32(S.BOUND) * np.sign(x[:8])
    return x


def host(t):
    return t.detach().float().cpu().numpy()


def close(a, b, rtol, atol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), f"{what}: worst excess {err.max():.3e}, max abs err {np.abs(a - b).max():.3e}"


BLOCKS = [("plane_static", slice(0, 32)), ("plane_dynamic", slice(32, 64)),
          ("hash_static", slice(64, 96)), ("hash_dynamic", slice(96, 120))]


@pytest.mark.parametrize("ti", range(len(TIMES)))
@pytest.mark.parametrize("lidar", [True, False])
def test_encoder_features_exact_positions(pkg, ti, lidar):
    """Zero flow (last flow layer = 0): all three queries sit at x, so every block of the
    120-vector (K-planes static/dynamic, hash static/dynamic, blended over t, t1, t2) must agree
    with the oracle to the resolution of the fp16 feature tile."""
    from oracle.field_oracle import FieldOracle
    p = dict(FC.oracle_params())
    p["flow_mlp"] = p["flow_mlp"].clone()
    p["flow_mlp"][-6 * 64:] = 0
    m = pkg.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND).eval()
    m.load_flat_params(p)
    x = pts(512, 7)
    t = TIMES[ti]
    with torch.no_grad():
        ef, eflow = FieldOracle(FC.oracle_config(), p).features(torch.from_numpy(x), t, lidar)
    gf, gflow = m.features(torch.from_numpy(x).cuda(), torch.tensor([[t]], device="cuda"), lidar)
    assert not host(gflow).any() and not eflow.numpy().any()
    ef, gf = ef.numpy(), host(gf)
    for name, sl in BLOCKS:
        scale = np.abs(ef[:, sl]).max()
        close(gf[:, sl], ef[:, sl], 1e-3, 1e-3 * scale, f"{name} t={t}")


@pytest.mark.parametrize("ti", range(len(TIMES)))
@pytest.mark.parametrize("lidar", [True, False])
def test_encoder_features_and_flow(model, orc, ti, lidar):
    """With the flow on: the flow itself, and the blocks at a tolerance that allows for the
    fp16 flow MLP moving the warped queries (see oracle/field_init.py, conditioning note)."""
    x = pts(512, 7)
    t = TIMES[ti]
    with torch.no_grad():
        ef, eflow = orc.features(torch.from_numpy(x), t, lidar)
    gf, gflow = model.features(torch.from_numpy(x).cuda(), torch.tensor([[t]], device="cuda"), lidar)
    close(host(gflow), eflow.numpy(), 5e-3, 2e-5, "flow")
    assert np.abs(eflow.numpy()).max() > 1e-3
    ef, gf = ef.numpy(), host(gf)
    for name, sl in BLOCKS:
        scale = np.abs(ef[:, sl]).max()
        close(gf[:, sl], ef[:, sl], 3e-3, 5e-3 * scale, f"{name} t={t}")


@pytest.mark.parametrize("ti", range(len(TIMES)))
@pytest.mark.parametrize("lidar", [True, False])
def test_density_vs_oracle_and_golden(model, orc, ti, lidar):
    gold = np.load(os.path.join(GOLDEN, "field_ref.npz"))
    x = gold["x"]
    t = float(gold["times"][ti])
    r = model.density(torch.from_numpy(x).cuda(), torch.tensor([[t]], device="cuda"), lidar)
    assert r["sigma"].dtype == torch.float32 and r["sigma"].shape == (256,) and r["geo_feat"].shape == (256, 15)
    k = f"den_t{ti}_{'l' if lidar else 'c'}_"
    close(host(r["sigma"]), gold[k + "sigma"], 1e-2, 0, "sigma vs reference modules")
    close(host(r["geo_feat"]), gold[k + "geo"], 1e-2, 1e-2 * np.abs(gold[k + "geo"]).max(), "geo_feat")
    with torch.no_grad():
        o = orc.density(torch.from_numpy(x), t, lidar)
    close(host(r["sigma"]), o["sigma"].numpy(), 1e-2, 0, "sigma vs oracle")
    f = model.flow(torch.from_numpy(x).cuda(), t)   # python float time is accepted too
    got = np.concatenate([host(f["flow_forward"]), host(f["flow_backward"])], -1)
    close(got, gold[f"flow_t{ti}"], 5e-3, 2e-5, "flow vs reference modules")


@pytest.mark.parametrize("ds", [1, 60])
@pytest.mark.parametrize("lidar", [True, False])
@pytest.mark.parametrize("perturb", [0, 1])
def test_run_vs_reference_golden(pkg, ds, lidar, perturb):
    gold = np.load(os.path.join(GOLDEN, "field_ref.npz"))
    m = pkg.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                        min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH,
                        density_scale=float(ds)).eval()
    m.load_flat_params(FC.oracle_params())
    k = f"ds{ds}_run_{'l' if lidar else 'c'}{perturb}_"
    o, d = torch.from_numpy(gold[k + "o"]).cuda()[None], torch.from_numpy(gold[k + "d"]).cuda()[None]
    noise = torch.from_numpy(gold[k + "noise"]).cuda() if perturb else None
    r = m.render(o, d, torch.tensor([[0.3]], device="cuda"), cal_lidar_color=lidar, staged=False, num_steps=40,
                 perturb=bool(perturb), noise=noise)
    sfx = "_lidar" if lidar else ""
    assert r["depth" + sfx].shape == (1, 48) and r["image" + sfx].shape == (1, 48, 2 if lidar else 3)
    close(host(r["z_vals"]), gold[k + "z_vals"], 1e-6, 1e-7, "z_vals")
    close(host(r["weights"]), gold[k + "weights"], 1e-2, 1e-5, "weights")
    close(host(r["weights_sum" + sfx]), gold[k + "weights_sum"], 1e-2, 1e-5, "weights_sum")
    close(host(r["depth" + sfx]).reshape(-1), gold[k + "depth"], 1e-2, 1e-5, "depth")
    close(host(r["image" + sfx]).reshape(48, -1), gold[k + "image"], 1e-2, 1e-4, "image")


def test_full_lidar_frame_properties(model, orc):
    """66x1030 frame, 128 samples: staged render == run on sub-batches; spot rays match the oracle."""
    o, d = S.lidar_rays(-1, seed=0)
    to, td = torch.from_numpy(o).cuda()[None], torch.from_numpy(d).cuda()[None]
    t = torch.tensor([[0.5]], device="cuda")
    full = model.render(to, td, t, cal_lidar_color=True, staged=True, num_steps=128)
    assert set(full) == {"depth_lidar", "image_lidar"} and full["depth_lidar"].shape == (1, 67980)
    img = host(full["image_lidar"])[0]
    assert np.isfinite(img).all() and (img >= 0).all() and (img <= 1).all()
    with torch.no_grad():   # same (inference) kernels as the staged render; with autograd on, run() takes the training forward
        sub = model.render(to[:, 5000:9096], td[:, 5000:9096], t, cal_lidar_color=True, staged=False, num_steps=128)
    np.testing.assert_allclose(host(sub["depth_lidar"]), host(full["depth_lidar"])[:, 5000:9096], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(host(sub["image_lidar"]), host(full["image_lidar"])[:, 5000:9096], rtol=1e-6, atol=1e-7)
    ws = host(sub["weights_sum_lidar"]); w = host(sub["weights"])
    np.testing.assert_allclose(w.sum(1), ws, rtol=1e-5)
    assert (ws <= 1 + 1e-5).all()
    idx = np.arange(0, 67980, 997)
    with torch.no_grad():
        e = orc.run(torch.from_numpy(o[idx]), torch.from_numpy(d[idx]), 0.5, True, 128)
    close(host(full["depth_lidar"])[0, idx], e["depth"].numpy(), 1e-2, 1e-5, "depth")
    close(img[idx], e["image"].numpy(), 1e-2, 1e-4, "image")


def test_params_repack_on_update(model):
    x = torch.from_numpy(pts(64, 3)).cuda()
    a = model.density(x, 0.4, True)["sigma"].clone()
    with torch.no_grad():
        model.sigma_net.mul_(1.5)
    b = model.density(x, 0.4, True)["sigma"]
    assert not torch.allclose(a, b)
    with torch.no_grad():
        model.sigma_net.div_(1.5)
    c = model.density(x, 0.4, True)["sigma"]
    torch.testing.assert_close(a, c, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("t", [0.37, 0.0, 1.0])
def test_density_modes_agree(pkg, model, t):
    """The staged density evaluation (mode 1), its variant with the dynamic 2-D hash tables staged
    in shared memory (mode 2: fp16 table copies, k_dyn_stage) and the single fused kernel (mode 0)
    run the same arithmetic: features / sigma / geo / rendered outputs agree to fp16 rounding.
    t = 0 and t = 1 exercise the first/last-frame aliasing of the warped queries."""
    L = pkg._lib.lib()
    x = torch.from_numpy(pts(40000, 11)).cuda()
    o, d = S.lidar_rays(300, seed=4)
    to, td = torch.from_numpy(o).cuda()[None], torch.from_numpy(d).cuda()[None]
    res = {}
    prev = L.nvsf_density_mode_get()
    try:
        for mode in (0, 1, 2):
            assert L.nvsf_set_option(b"density_mode", mode) == 0
            den = model.density(x, t, True)
            feats, _ = model.features(x, t, True)
            r = model.render(to, td, torch.tensor([[t]], device="cuda"), cal_lidar_color=True, num_steps=96)
            res[mode] = (host(den["sigma"]), host(den["geo_feat"]), host(r["depth_lidar"]), host(r["image_lidar"]),
                         host(feats))
    finally:
        L.nvsf_set_option(b"density_mode", prev)
    assert L.nvsf_set_option(b"density_mode", 7) == -1 and L.nvsf_set_option(b"nope", 0) == -1
    for mode in (0, 2):
        for a, b, name in zip(res[mode], res[1], ("sigma", "geo", "depth", "image", "features")):
            close(a, b, 2e-3, 2e-3 * np.abs(b).max(), f"mode {mode} {name}")
    # the static hash block does not depend on the mode at all (mode 2 reads the planes from fp16 texels)
    assert np.array_equal(res[2][4][:, 64:96], res[1][4][:, 64:96])


def test_dyn_stage_tiles_and_chunks(pkg, model):
    """Mode 2 with small work items and chunks (several tiles per table type, several chunks per
    call, a ragged tail) equals mode 2 with the defaults bit for bit, for explicit points and for
    the ray path."""
    L = pkg._lib.lib()
    x = torch.from_numpy(pts(150001, 5)).cuda()
    o, d = S.lidar_rays(700, seed=9)
    to, td = torch.from_numpy(o).cuda()[None], torch.from_numpy(d).cuda()[None]
    prev = L.nvsf_density_mode_get()
    out = []
    try:
        assert L.nvsf_set_option(b"density_mode", 2) == 0
        for tile, chunk in ((32768, 1024), (1024, 1)):
            assert L.nvsf_set_option(b"dyn_tile", tile) == 0 and L.nvsf_set_option(b"split_chunk", chunk) == 0
            den = model.density(x, 0.6, False)
            r = model.render(to, td, torch.tensor([[0.6]], device="cuda"), cal_lidar_color=True, num_steps=128)
            out.append((host(den["sigma"]), host(den["geo_feat"]), host(r["depth_lidar"]), host(r["image_lidar"])))
    finally:
        L.nvsf_set_option(b"dyn_tile", 32768); L.nvsf_set_option(b"split_chunk", 1024)
        L.nvsf_set_option(b"density_mode", prev)
    for a, b in zip(*out):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n", [128 * 37, 100001])
def test_sigma_stage_tcgen05_matches_mma_sync(pkg, model, n):
    """The sigma MLP on tcgen05.mma / TMEM (csrc/sigma_tc.cu, option sigma_tc) against the mma.sync
    stage: same fp16 operands, fp32 accumulation in a different order -> agreement to fp16 rounding
    of the hidden layer; covers full tiles and a ragged tail."""
    L = pkg._lib.lib()
    x = torch.from_numpy(pts(n, 21)).cuda()
    out = {}
    try:
        for tc in (0, 1):
            assert L.nvsf_set_option(b"sigma_tc", tc) == 0
            den = model.density(x, 0.45, True)
            torch.cuda.synchronize()
            out[tc] = (host(den["sigma"]), host(den["geo_feat"]))
    finally:
        L.nvsf_set_option(b"sigma_tc", 1)
    close(out[1][0], out[0][0], 2e-3, 0, "sigma tcgen05 vs mma.sync")
    close(out[1][1], out[0][1], 2e-3, 2e-3 * np.abs(out[0][1]).max(), "geo tcgen05 vs mma.sync")


@pytest.mark.parametrize("n", [1024 * 9, 100001])
def test_fused_gather_sigma_tcgen05_matches_staged(pkg, model, n):
    """Mode 2 with the gather stage fused with the sigma MLP (k_encode_sigma_tc: feature rows written
    straight into the swizzled UMMA operand tile, two K halves accumulated in TMEM) against the staged
    pair k_encode_stage -> k_sigma_stage_tc: identical features, so sigma / geo / rendered outputs agree
    to the rounding of the fp32 accumulation order."""
    L = pkg._lib.lib()
    x = torch.from_numpy(pts(n, 23)).cuda()
    o, d = S.lidar_rays(500, seed=12)
    to, td = torch.from_numpy(o).cuda()[None], torch.from_numpy(d).cuda()[None]
    out = {}
    try:
        for fuse in (0, 1):
            assert L.nvsf_set_option(b"fuse_sigma", fuse) == 0
            den = model.density(x, 0.45, False)
            with torch.no_grad():
                r = model.render(to, td, torch.tensor([[0.45]], device="cuda"), cal_lidar_color=True, num_steps=96)
            torch.cuda.synchronize()
            out[fuse] = (host(den["sigma"]), host(den["geo_feat"]), host(r["depth_lidar"]), host(r["image_lidar"]))
    finally:
        L.nvsf_set_option(b"fuse_sigma", 1)
    for a, b, name in zip(out[1], out[0], ("sigma", "geo", "depth", "image")):
        close(a, b, 1e-3, 1e-3 * np.abs(b).max(), f"fused vs staged {name}")


@pytest.mark.parametrize("t", [0.45, 0.0])
def test_flow_stage_tcgen05_matches_mma_sync(pkg, model, t):
    """Flow stage on tcgen05 (k_flow_tc: flow-grid gather into the UMMA operand tile, three chained
    MMA batches through TMEM) against the mma.sync flow stage: flow, and everything downstream."""
    L = pkg._lib.lib()
    x = torch.from_numpy(pts(70001, 29)).cuda()
    out = {}
    try:
        for tc in (0, 1):
            assert L.nvsf_set_option(b"flow_tc", tc) == 0
            f = model.flow(x, t)
            den = model.density(x, t, True)
            torch.cuda.synchronize()
            out[tc] = (np.concatenate([host(f["flow_forward"]), host(f["flow_backward"])], -1),
                       host(den["sigma"]), host(den["geo_feat"]))
        # option flow_ts: the hidden activations stay in tensor memory (second layer as two N = 32 halves): the
        # same products and fp16 roundings as the shared-memory form -> bit-identical
        assert L.nvsf_set_option(b"flow_ts", 1) == 0 and L.nvsf_get_option(b"flow_ts") == 1
        f = model.flow(x, t)
        den = model.density(x, t, True)
        torch.cuda.synchronize()
        out[2] = (np.concatenate([host(f["flow_forward"]), host(f["flow_backward"])], -1),
                  host(den["sigma"]), host(den["geo_feat"]))
    finally:
        L.nvsf_set_option(b"flow_tc", 1)
        L.nvsf_set_option(b"flow_ts", 0)
    assert np.abs(out[0][0]).max() > 1e-4
    close(out[1][0], out[0][0], 2e-3, 2e-3 * np.abs(out[0][0]).max(), "flow tcgen05 vs mma.sync")
    close(out[1][1], out[0][1], 2e-3, 0, "sigma")
    close(out[1][2], out[0][2], 2e-3, 2e-3 * np.abs(out[0][2]).max(), "geo")
    for a, b in zip(out[2], out[1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n", [1, 7, 127, 129, 1025])
def test_tiny_and_ragged_point_counts(pkg, model, orc, n):
    """Counts below and just above one 128-sample UMMA tile / one 1024-thread CTA through the
    default tcgen05 pipeline: every sample equals the same sample evaluated inside a large batch,
    and the oracle within the field tolerance."""
    x = pts(2048, 31)
    big = model.density(torch.from_numpy(x).cuda(), 0.37, True)
    small = model.density(torch.from_numpy(x[:n]).cuda(), 0.37, True)
    assert small["sigma"].shape == (n,) and small["geo_feat"].shape == (n, 15)
    assert torch.equal(small["sigma"], big["sigma"][:n]) and torch.equal(small["geo_feat"], big["geo_feat"][:n])
    with torch.no_grad():
        o = orc.density(torch.from_numpy(x[:n]), 0.37, True)
    close(host(small["sigma"]), o["sigma"].numpy(), 1e-2, 0, "sigma vs oracle")


def test_empty_point_set(model):
    r = model.density(torch.empty(0, 3, device="cuda"), 0.5, True)
    assert r["sigma"].shape == (0,) and r["geo_feat"].shape == (0, 15)


@pytest.mark.parametrize("lidar", [True, False])
@pytest.mark.parametrize("steps", [768, 200, 128, 37])
def test_composite_heads_tcgen05_matches_mma_sync(pkg, model, lidar, steps):
    """Compositing with the head MLPs on tcgen05.mma / TMEM (csrc/render.cu k_composite_tc, option
    heads_tc: a warpgroup per ray, 128-sample UMMA tiles, three chained MMA batches per tile) against
    the mma.sync renderer k_render_composite: same fp16 operands and masks; the transmittance scan is
    128 instead of 32 samples wide and the fp32 sums run in a different order.  Covers whole tiles,
    ragged last tiles (200, 37 samples), both modalities, weights / z_vals outputs and perturbation."""
    L = pkg._lib.lib()
    if lidar:
        o, d = S.lidar_rays(333, seed=41)
    else:
        o, d = S.camera_rays(333, seed=42)
    to, td = torch.from_numpy(o).cuda()[None], torch.from_numpy(d).cuda()[None]
    noise = torch.rand(333, steps, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    out = {}
    try:
        for tc in (0, 1, 2, 3, 4, 5, 6):   # 6: hidden activations as the A operand from tensor memory; 2..5: eight / six / five / seven warpgroups per CTA, the nets one after the
            # other, the per-ray direction term added by a second layer-1 MMA (fp16 u) instead of the epilogue
            assert L.nvsf_set_option(b"heads_tc", tc) == 0
            assert L.nvsf_get_option(b"heads_tc") == tc
            with torch.no_grad():
                r = model.run(to, td, torch.tensor([[0.45]], device="cuda"), cal_lidar_color=lidar, num_steps=steps,
                              perturb=True, noise=noise)
            torch.cuda.synchronize()
            sfx = "_lidar" if lidar else ""
            out[tc] = (host(r["depth" + sfx]), host(r["image" + sfx]), host(r["weights_sum" + sfx]),
                       host(r["weights"]), host(r["z_vals"]))
    finally:
        L.nvsf_set_option(b"heads_tc", 6)
    assert np.abs(out[0][1]).max() > 1e-3
    for tc in (1, 2, 3, 4, 5, 6):
        assert np.array_equal(out[tc][4], out[0][4])
        for a, b, name in zip(out[tc], out[0], ("depth", "image", "weights_sum", "weights")):
            close(a, b, 1e-4 if name != "image" else 2e-3, 1e-6 if name != "image" else 2e-3 * np.abs(b).max(),
                  f"heads tcgen05 (mode {tc}) vs mma.sync {name} (S={steps})")
    # the same tile arithmetic, only the number of warpgroups differs: these forms agree bit for bit
    for tc in (3, 4, 5):
        for a, b in zip(out[tc], out[2]):
            assert np.array_equal(a, b)
    for a, b in zip(out[6], out[2]):   # ... and so does the form that keeps the activations in tensor memory
        assert np.array_equal(a, b)


@pytest.mark.parametrize("t", [0.45, 1.0])
def test_fused_stage_packed_half_interpolation(pkg, model, orc, t):
    """k_encode_sigma_tc<true> (option half_math, off by default): plane / static-hash interpolation on HFMA2
    (tcnn interpolates its grids in half precision too) against the same kernel with fp32
    interpolation, and against the CPU oracle at the north star's tolerance for MLP outputs (1e-2)."""
    L = pkg._lib.lib()
    x = pts(50001, 37)
    xd = torch.from_numpy(x).cuda()
    o, d = S.lidar_rays(400, seed=14)
    to, td = torch.from_numpy(o).cuda()[None], torch.from_numpy(d).cuda()[None]
    out = {}
    try:
        for hm in (0, 1):
            assert L.nvsf_set_option(b"half_math", hm) == 0
            den = model.density(xd, t, True)
            with torch.no_grad():
                r = model.render(to, td, torch.tensor([[t]], device="cuda"), cal_lidar_color=True, num_steps=256)
            torch.cuda.synchronize()
            out[hm] = (host(den["sigma"]), host(den["geo_feat"]), host(r["depth_lidar"]), host(r["image_lidar"]))
    finally:
        L.nvsf_set_option(b"half_math", 0)
    gmax = np.abs(out[0][1]).max()
    close(out[1][0], out[0][0], 5e-3, 0, "sigma half vs fp32 interpolation")
    close(out[1][1], out[0][1], 5e-3, 5e-3 * gmax, "geo half vs fp32 interpolation")
    close(out[1][2], out[0][2], 2e-3, 0, "depth half vs fp32 interpolation")
    close(out[1][3], out[0][3], 2e-3, 2e-3 * np.abs(out[0][3]).max(), "image half vs fp32 interpolation")
    with torch.no_grad():
        ref = orc.density(torch.from_numpy(x[:4096]), t, True)
    close(out[1][0][:4096], ref["sigma"].numpy(), 1e-2, 0, "sigma (half interpolation) vs oracle")


def test_whole_frame_chunk_vs_feature_scratch_chunks(pkg, model):
    """More samples than the [4 M,128] feature scratch of the un-fused path holds: the fused tcgen05
    gather stage takes them as ONE chunk (default split_chunk = 64 M samples), the un-fused path
    (fuse_sigma = 0) must fall back to 4 M-sample chunks, and small explicit chunks must give the same
    rows — every sample is independent of how the batch is cut."""
    L = pkg._lib.lib()
    n = (4 << 20) + 70001
    rng = torch.Generator(device="cuda").manual_seed(17)
    x = (torch.rand(n, 3, device="cuda", generator=rng) * 2 - 1) * 1.9
    out = {}
    try:
        for name, fuse, chunk in (("one_chunk", 1, 1024), ("unfused", 0, 1024), ("small_chunks", 1, 16)):
            assert L.nvsf_set_option(b"fuse_sigma", fuse) == 0 and L.nvsf_set_option(b"split_chunk", chunk) == 0
            den = model.density(x, 0.45, True)
            torch.cuda.synchronize()
            out[name] = (den["sigma"], den["geo_feat"])
    finally:
        L.nvsf_set_option(b"fuse_sigma", 1)
        L.nvsf_set_option(b"split_chunk", 1024)
    assert torch.equal(out["one_chunk"][0], out["small_chunks"][0]) and torch.equal(out["one_chunk"][1], out["small_chunks"][1])
    close(host(out["one_chunk"][0]), host(out["unfused"][0]), 1e-3, 0, "sigma fused one chunk vs un-fused 4 M chunks")
    g = host(out["unfused"][1])
    close(host(out["one_chunk"][1]), g, 1e-3, 1e-3 * np.abs(g).max(), "geo fused one chunk vs un-fused 4 M chunks")


def test_options_are_scoped_to_the_model(pkg, model):
    """NeRFNetwork(options=...) applies its overrides only around its own launches: same bits as the process-wide switch,
    the process defaults untouched afterwards, another model in the same process unaffected."""
    L = pkg._lib.lib()
    o, d = S.lidar_rays(97, seed=3)
    to, td = torch.from_numpy(o).cuda()[None], torch.from_numpy(d).cuda()[None]
    t = torch.tensor([[0.45]], device="cuda")
    other = pkg.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                            min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH,
                            options={"heads_tc": 0, "dyn_tile": 8192})
    other.load_flat_params(FC.oracle_params())
    other.eval()
    base = {k: L.nvsf_get_option(k) for k in (b"heads_tc", b"dyn_tile")}

    def run(m):
        with torch.no_grad():
            r = m.run(to, td, t, cal_lidar_color=True, num_steps=200)
        return host(r["depth_lidar"]), host(r["image_lidar"]), host(r["weights"])

    r_def = run(model)
    r_other = run(other)
    assert {k: L.nvsf_get_option(k) for k in base} == base
    assert all(np.array_equal(a, b) for a, b in zip(run(model), r_def))
    try:
        assert L.nvsf_set_option(b"heads_tc", 0) == 0 and L.nvsf_set_option(b"dyn_tile", 8192) == 0
        r_glob = run(model)
    finally:
        for k, v in base.items():
            L.nvsf_set_option(k, v)
    assert all(np.array_equal(a, b) for a, b in zip(r_other, r_glob))
    close(r_other[1], r_def[1], 2e-3, 2e-3 * np.abs(r_def[1]).max(), "image, mma.sync heads (model-scoped) vs defaults")
