"""Diagnostic: per-level relative error of the flow-grid gradient of the scene-flow loss, against the oracle with and
without the fp16 cast of its table gradients."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import field_cases as FC
import test_loss_terms_gpu as TL
from oracle import loss_oracle as LO
from oracle import tcnn_standin as T
from oracle.field_oracle import FieldOracle
pkg = importlib.import_module("selfsupervised-nvsf_b200")
rng = np.random.default_rng(3); M = 1500
pc = ((rng.random((M, 3), dtype=np.float32) * 2 - 1) * np.float32(0.8)).astype(np.float32)
pcf = (pc + 0.01 * rng.standard_normal((M + 17, 3)).astype(np.float32)[:M]).astype(np.float32)
pcb = (pc[::2] - 0.01).astype(np.float32)
base = FC.oracle_params()
cfg = FC.oracle_config()
tt = torch.from_numpy


def oracle_grad(exact):
    leaf = {k: base[k].clone().requires_grad_(True) for k in ("flow_grid", "flow_mlp")}
    orc = FieldOracle(cfg, dict(base, **leaf))
    old = T._fp16_round
    if exact:   # table values still fp16, gradient not cast
        class R(torch.autograd.Function):
            @staticmethod
            def forward(ctx, p): return p.to(torch.float16).to(torch.float32)
            @staticmethod
            def backward(ctx, g): return g
        T._fp16_round = R.apply
    try:
        (LO.flow_loss(lambda x: orc.flow(x, 0.4), TL._cham, tt(pc), tt(pcf), tt(pcb)) * FC.LOSS_SCALE).backward()
    finally:
        T._fp16_round = old
    return leaf["flow_grid"].grad.numpy().astype(np.float64) / FC.LOSS_SCALE


g16, g32 = oracle_grad(False), oracle_grad(True)
m = TL._model(pkg)
cu = lambda a: torch.from_numpy(a).cuda()
pkg.losses.flow_loss(m, cu(pc), torch.tensor([[0.4]], device="cuda"), cu(pcf), cu(pcb)).backward()
got = m.flow_grid.grad.detach().cpu().numpy().reshape(-1).astype(np.float64)
rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
print("ours vs oracle(fp16 grad cast)", rel(got, g16), " ours vs oracle(fp32 grads)", rel(got, g32), " oracle16 vs oracle32", rel(g16, g32))
for l, lv in enumerate(cfg.flow_grid_levels):
    a, b = lv["offset"] * 8, (lv["offset"] + lv["size"]) * 8
    print(f"level {l:2d} res {lv['res']:5d} hashed {int(lv['hashed'])} |g| {np.linalg.norm(g32[a:b]):.3e} ours-vs-32 {rel(got[a:b], g32[a:b]):.4f} 16-vs-32 {rel(g16[a:b], g32[a:b]):.4f}")
# per-entry distribution: a systematic error would move the median, flipped ReLUs leave it alone and sit in the tail
nz = np.abs(g32) > 1e-3 * np.abs(g32).max()
r = np.abs(got[nz] - g32[nz]) / np.abs(g32[nz])
print("touched entries", int(nz.sum()), "median rel err", np.median(r), "p75", np.percentile(r, 75), "p90", np.percentile(r, 90),
      "p99", np.percentile(r, 99), "frac > 1e-2", float((r > 1e-2).mean()), "frac > 5e-2", float((r > 5e-2).mean()))
r16 = np.abs(g16[nz] - g32[nz]) / np.abs(g32[nz])
print("oracle fp16-cast vs fp32: median", np.median(r16), "p90", np.percentile(r16, 90), "frac > 1e-2", float((r16 > 1e-2).mean()))
