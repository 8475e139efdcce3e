"""TEST INFRASTRUCTURE.  Generates tests/golden/raymarching_ref_sm100a.npz by running the
reference's OWN extension (oracle/_ref/_raymarching_ref.so = unmodified
nvsf/nerf/raymarching/src/raymarching.cu built for sm_100a by oracle/build_ref.sh) on seeded
inputs.  Needs a GPU:   gpurun -- python oracle/make_golden_raymarching.py gpurun_out/golden
then copy the .npz into tests/golden/.  The CPU test suite pins the C oracle against these
vectors (tests/test_oracle_golden.py)."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

S = cases.S
C, H, BOUND = S.CASCADE, S.GRID_SIZE, S.BOUND


def load_ref():
    path = os.path.join(ROOT, "oracle", "_ref", "_raymarching_ref.so")
    spec = importlib.util.spec_from_file_location("_raymarching_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def main(out_dir):
    ref = load_ref()
    out = {}
    # ---- near_far
    o, d = S.camera_rays(512, seed=41)
    o = o.copy(); d = d.copy()
    o[::7] *= 9.0; d[5::11, 0] = 0.0; d[3::13] *= -1.0
    n, f = torch.empty(512, device="cuda"), torch.empty(512, device="cuda")
    ref.near_far_from_aabb(dev(o), dev(d), dev(S.AABB), 512, S.MIN_NEAR, n, f)
    out.update(nf_o=o, nf_d=d, nf_nears=host(n), nf_fars=host(f))
    # ---- sph
    o2, d2 = S.camera_rays(256, seed=42)
    c = torch.empty(256, 2, device="cuda")
    ref.sph_from_ray(dev(o2), dev(d2), 3.0, 256, c)
    out.update(sph_o=o2, sph_d=d2, sph_coords=host(c))
    # ---- morton
    rng = np.random.default_rng(43)
    coords = np.concatenate([rng.integers(0, 128, size=(600, 3)), rng.integers(-2**31, 2**31 - 1, size=(200, 3))]).astype(np.int32)
    idx = torch.empty(800, dtype=torch.int32, device="cuda")
    ref.morton3D(dev(coords), 800, idx)
    back = torch.empty(800, 3, dtype=torch.int32, device="cuda")
    ref.morton3D_invert(idx, 800, back)
    out.update(mt_coords=coords, mt_indices=host(idx), mt_back=host(back))
    # ---- packbits
    g = rng.normal(size=(1, 8 * 1000)).astype(np.float32); g[0, ::5] = 0.25
    bfo = torch.empty(1000, dtype=torch.uint8, device="cuda")
    ref.packbits(dev(g), 1000, 0.25, bfo)
    out.update(pb_grid=g, pb_bits=host(bfo))
    # ---- march_rays_train + composite_rays_train
    N, max_steps = 96, 128
    for ci, (kind, fill, perturb, dt_gamma) in enumerate([
            ("lidar", "shell", False, S.DT_GAMMA), ("lidar", "random5", True, 0.0),
            ("camera", "full", True, S.DT_GAMMA), ("camera", "shell", True, S.DT_GAMMA)]):
        o, d, nears, fars, noises = cases.march_inputs(kind, N, seed=50 + ci, perturb=perturb)
        bf = cases.bitfield(fill, seed=6)
        M = N * max_steps
        xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
        rays = torch.empty(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        ref.march_rays_train(dev(o), dev(d), dev(bf), BOUND, dt_gamma, max_steps, N, C, H, M, dev(nears), dev(fars),
                             xyzs, dirs, deltas, rays, counter, dev(noises))
        torch.cuda.synchronize()
        m = int(host(counter)[0])
        cr, (cx, cd, cl) = cases.canonical_from_rays(host(rays), [host(xyzs), host(dirs), host(deltas)])
        sig, rgb = cases.field_values(m, seed=60 + ci)
        sig *= np.float32(30.0)
        ws = torch.empty(N, device="cuda"); de = torch.empty(N, device="cuda"); im = torch.empty(N, 3, device="cuda")
        ref.composite_rays_train_forward(dev(sig), dev(rgb), dev(cl), dev(cr), m, N, 1e-4, ws, de, im)
        g_ws = rng.normal(size=N).astype(np.float32); g_im = rng.normal(size=(N, 3)).astype(np.float32)
        gs = torch.zeros(m, device="cuda"); gr = torch.zeros(m, 3, device="cuda")
        ref.composite_rays_train_backward(dev(g_ws), dev(g_im), dev(sig), dev(rgb), dev(cl), dev(cr), ws, im, m, N, 1e-4, gs, gr)
        p = f"mt{ci}_"
        out.update({p + "meta": np.array([ci, N, max_steps, perturb, m], np.int64), p + "kind": np.array(kind), p + "fill": np.array(fill),
                    p + "dt_gamma": np.float32(dt_gamma), p + "o": o, p + "d": d, p + "nears": nears, p + "fars": fars, p + "noises": noises,
                    p + "rays": cr, p + "xyzs": cx, p + "deltas": cl, p + "sig_seed": np.int64(60 + ci),
                    p + "ws": host(ws), p + "depth": host(de), p + "image": host(im), p + "g_ws": g_ws, p + "g_im": g_im,
                    p + "grad_sigmas": host(gs), p + "grad_rgbs": host(gr)})
    # ---- one inference step (march_rays + composite_rays)
    N = 128
    o, d, nears, fars, _ = cases.march_inputs("lidar", N, seed=70, perturb=False)
    bf = cases.bitfield("shell", seed=6)
    n_alive, n_step = 100, 6
    alive = rng.permutation(N)[:n_alive].astype(np.int32)
    rays_t = (nears + rng.random(N, dtype=np.float32) * 0.2).astype(np.float32)
    noises = rng.random(n_alive, dtype=np.float32)
    M = n_alive * n_step
    xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
    ref.march_rays(n_alive, n_step, dev(alive), dev(rays_t), dev(o), dev(d), BOUND, S.DT_GAMMA, 1024, C, H, dev(bf),
                   dev(nears), dev(fars), xyzs, dirs, deltas, dev(noises))
    sig, rgb = cases.field_values(M, seed=71); sig *= np.float32(60.0)
    ws0 = rng.random(N, dtype=np.float32) * 0.5; de0 = rng.random(N, dtype=np.float32); im0 = rng.random((N, 3), dtype=np.float32)
    a_t, t_t, ws_t, de_t, im_t = dev(alive), dev(rays_t), dev(ws0), dev(de0), dev(im0)
    ref.composite_rays(n_alive, n_step, 1e-2, a_t, t_t, dev(sig), dev(rgb), deltas, ws_t, de_t, im_t)
    out.update(inf_o=o, inf_d=d, inf_nears=nears, inf_fars=fars, inf_alive=alive, inf_rays_t=rays_t, inf_noises=noises,
               inf_meta=np.array([n_alive, n_step], np.int64), inf_xyzs=host(xyzs), inf_dirs=host(dirs), inf_deltas=host(deltas),
               inf_ws0=ws0, inf_de0=de0, inf_im0=im0, inf_alive_out=host(a_t), inf_t_out=host(t_t), inf_ws=host(ws_t),
               inf_depth=host(de_t), inf_image=host(im_t))
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "raymarching_ref_sm100a.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", torch.cuda.get_device_name(0))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
