"""Diagnostic (GPU box): compare the CUDA backward's per-sample intermediates (d sigma-net output,
d features, d flow) with the CPU oracle's autograd for one gradient case."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
import field_cases as FC
import test_field_grad_gpu as T
from oracle.field_oracle import FieldOracle, mlp, _TruncExp
from oracle import raymarching_oracle as RO
pkg = importlib.import_module("selfsupervised-nvsf_b200")
F = importlib.import_module("selfsupervised-nvsf_b200.field")
gold = np.load(os.path.join(ROOT, "tests", "golden", "field_grad_ref.npz"))
tag = sys.argv[1] if len(sys.argv) > 1 else "c_last"
case = FC.grad_case(gold, tag)
lidar = case["lidar"]
m = T.make_model(pkg, case["ds"]); m._debug_keep = True
loss, out = T.run_case(m, case); loss.backward(); torch.cuda.synchronize()
N, S = m._debug["N"], m._debug["S"]; n = N * S
lay = (ctypes.c_size_t * 12)()
F._setup_lib().nvsf_render_uniform_debug_layout(ctypes.byref(m._cfg), N, S, lay)
sc, sv = m._debug["scratch"], m._debug["saved"]
def view(buf, off, cnt, dt): return buf[off:off + cnt * torch.tensor([], dtype=dt).element_size()].view(dt).cpu().float().numpy()
dgeo16 = view(sc, lay[6], n * 16, torch.float32).reshape(n, 16)
dfeat = view(sc, lay[7], n * 128, torch.float32).reshape(n, 128)[:, :120]
dflow = view(sc, lay[8], n * 8, torch.float32).reshape(n, 8)[:, :6]
feats_c = view(sv, lay[3], n * 128, torch.float16).reshape(n, 128)[:, :120]
geo_c = view(sv, lay[1], n * 16, torch.float16).reshape(n, 16)
# ---- oracle with hooks
p = FC.oracle_params()
orc = FieldOracle(FC.oracle_config(density_scale=case["ds"]), p)
c = orc.cfg
o, d = torch.from_numpy(case["o"]), torch.from_numpy(case["d"])
if lidar:
    nears = torch.full((N,), c.min_near_lidar); fars = torch.full((N,), c.lidar_max_depth)
else:
    a, b = RO.near_far_from_aabb(case["o"], case["d"], FC.S.AABB, FC.S.MIN_NEAR); nears, fars = torch.from_numpy(a), torch.from_numpy(b)
keep = {}
orig_features = orc.features
def features(x, t, lid):
    xn = (x + c.bound) / (2 * c.bound)
    flow = orc.flow_net(xn, t).detach().requires_grad_(True)
    keep["flow"] = flow
    orc.flow_net = lambda xn_, t_: flow
    f, _ = orig_features(x, t, lid)
    f = f  # graph goes through `flow` leaf for the warped planes
    f.retain_grad(); keep["feats"] = f
    return f, flow
orc.features = features
orig_density = orc.density
def density(x, t, lid):
    feats, _ = orc.features(x, t, lid)
    h = mlp(p["sigma_net"], [(64, 128), (16, 64)], feats, 16)
    h.retain_grad(); keep["h"] = h
    return {"sigma": _TruncExp.apply(h[:, 0]), "geo_feat": h[:, 1:]}
orc.density = density
r = orc.run(o, d, case["t"], lidar, S, nears, fars, None if case["noise"] is None else torch.from_numpy(case["noise"]))
l = FC.linear_loss(r, case["coef"]); l.backward()
dh, df, dfl = keep["h"].grad.numpy(), keep["feats"].grad.numpy(), keep["flow"].grad.numpy()
def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
print(tag, "loss", loss.item(), l.item())
print("fwd feats rel", rel(feats_c, keep["feats"].detach().numpy()), "h rel", rel(geo_c, keep["h"].detach().numpy()))
print("d logit  rel", rel(dgeo16[:, 0], dh[:, 0]), " norm", np.linalg.norm(dh[:, 0]))
print("d geo    rel", rel(dgeo16[:, 1:], dh[:, 1:]), " norm", np.linalg.norm(dh[:, 1:]))
print("d feats  rel", rel(dfeat, df), " per block", [round(float(rel(dfeat[:, a:b], df[:, a:b])), 4) for a, b in ((0, 32), (32, 64), (64, 96), (96, 120))])
print("d flow   rel", rel(dflow, dfl), " norm", np.linalg.norm(dfl))
# worst samples of d logit
e = np.abs(dgeo16[:, 0] - dh[:, 0]); k = np.argsort(-e)[:8]
for i in k: print("  sample", i, "ray", i // S, "k", i % S, "cuda", dgeo16[i, 0], "oracle", dh[i, 0], "w", r["weights"].detach().numpy().reshape(-1)[i])
# worst rows of d geo
rg = view(sv, lay[5], n * 4, torch.float32).reshape(n, 4)
eg = np.linalg.norm(dgeo16[:, 1:] - dh[:, 1:], axis=1); k = np.argsort(-eg)[:10]
wv = r["weights"].detach().numpy().reshape(-1)
orc_rgb = None
for i in k:
    print("  dgeo row", i, "ray", i // S, "k", i % S, "err", eg[i], "|cuda|", np.linalg.norm(dgeo16[i, 1:]), "|oracle|", np.linalg.norm(dh[i, 1:]), "w", wv[i], "rgb", rg[i])
print("rows with oracle dgeo != 0:", int((np.abs(dh[:, 1:]).sum(1) > 0).sum()), " cuda:", int((np.abs(dgeo16[:, 1:]).sum(1) > 0).sum()))
oz = (np.abs(dh).sum(1) == 0); cz = (np.abs(dgeo16).sum(1) == 0)
print("rows: oracle zero", int(oz.sum()), "cuda zero", int(cz.sum()), "oracle zero & cuda nonzero", int((oz & ~cz).sum()))
for i in np.flatnonzero(oz & ~cz)[:12]:
    print("   row", i, "ray", i // S, "k", i % S, "cuda dlogit", dgeo16[i, 0], "|dgeo|", np.abs(dgeo16[i, 1:]).max(), "w", wv[i], "sigma", r["sigma"].detach().numpy().reshape(-1)[i])
ofz = (np.abs(df).sum(1) == 0); cfz = (np.abs(dfeat).sum(1) == 0)
print("dfeat rows: oracle zero", int(ofz.sum()), "cuda zero", int(cfz.sum()))
