"""Gradient error of the CUDA backward vs the CPU oracle's autograd for several (rays, samples) shapes — separates
the small-sample noise floor (fp16 ReLU flips that do not average out) from a shape-dependent bug."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
pkg = importlib.import_module("selfsupervised-nvsf_b200")
import field_cases as FC
S = FC.S
import test_field_grad_gpu as T
L = pkg._lib.lib()
for lidar, N, Sn in [(True, 37, 37), (True, 32, 32), (True, 64, 64), (True, 37, 64), (True, 150, 37), (True, 128, 128),
                     (False, 21, 150), (False, 32, 128)]:
    rng = np.random.default_rng(23)
    o, d = (S.lidar_rays if lidar else S.camera_rays)(N, seed=5)
    nch = 2 if lidar else 3
    case = dict(lidar=lidar, t=0.45, ds=1.0, o=o, d=d, noise=None,
                coef=dict(a=rng.normal(size=N).astype(np.float32), b=rng.normal(size=(N, nch)).astype(np.float32),
                          c=(0.1 * rng.normal(size=(N, Sn))).astype(np.float32), e=rng.normal(size=N).astype(np.float32)))
    e, eloss, _ = FC.oracle_grads(case)
    fl = FC.oracle_grad_floor(case, f"r{lidar}{N}{Sn}")
    for tc in (1, 0):
        L.nvsf_set_option(b"mlp_bwd_tc", tc)
        m = T.make_model(pkg, 1.0)
        loss, _ = T.run_case(m, case)
        loss.backward()
        g = T.grads_of(m, lidar)
        errs = {}
        for name in FC.GRAD_NAMES:
            ref = e[name].reshape(-1).astype(np.float64)
            if ref.any():
                errs[name] = f"{np.linalg.norm(g[name] - ref) / np.linalg.norm(ref):.4f}/{fl[name]:.4f}"
        print(lidar, N, Sn, "tc", tc, f"loss {float(loss):.5f} vs {eloss:.5f}", errs, flush=True)
L.nvsf_set_option(b"mlp_bwd_tc", 1)
