"""Experiment: gradient error vs the power-of-two scale shift of the fp16 MLP backward kernels."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import field_cases as FC
import test_field_grad_gpu as T
import test_loss_terms_gpu as TL
from oracle import loss_oracle as LO
from oracle.field_oracle import FieldOracle
pkg = importlib.import_module("selfsupervised-nvsf_b200")
L = pkg._lib.lib()
gold = np.load(os.path.join(ROOT, "tests", "golden", "field_grad_ref.npz"))
rng = np.random.default_rng(3); M = 1500
pc = ((rng.random((M, 3), dtype=np.float32) * 2 - 1) * np.float32(0.8)).astype(np.float32)
pcf = (pc + 0.01 * rng.standard_normal((M + 17, 3)).astype(np.float32)[:M]).astype(np.float32)
pcb = (pc[::2] - 0.01).astype(np.float32)
base = FC.oracle_params()
leaf = {k: base[k].clone().requires_grad_(True) for k in ("flow_grid", "flow_mlp")}
orc = FieldOracle(FC.oracle_config(), dict(base, **leaf))
tt = torch.from_numpy
LO.flow_loss(lambda x: orc.flow(x, 0.4), TL._cham, tt(pc), tt(pcf), tt(pcb)).backward()
refs = {t: FC.oracle_grads(FC.grad_case(gold, t))[0] for t in ("l_mid", "c_last")}
for sf, ss, sh in ((0, 0, 2), (3, 0, 2), (6, 0, 2), (6, 3, 2), (6, 5, 4), (8, 5, 5)):
    for k, v in ((b"bwd_shift_flow", sf), (b"bwd_shift_sigma", ss), (b"bwd_shift_heads", sh)):
        assert L.nvsf_set_option(k, v) == 0
    m = TL._model(pkg)
    cu = lambda a: torch.from_numpy(a).cuda()
    pkg.losses.flow_loss(m, cu(pc), torch.tensor([[0.4]], device="cuda"), cu(pcf), cu(pcb)).backward()
    row = [f"shift flow/sigma/heads {sf}/{ss}/{sh}: flow_loss"]
    for name in ("flow_grid", "flow_mlp"):
        got = getattr(m, name).grad.detach().cpu().numpy().reshape(-1).astype(np.float64)
        want = leaf[name].grad.numpy().reshape(-1).astype(np.float64)
        row.append(f"{name}={np.linalg.norm(got - want) / np.linalg.norm(want):.4f}")
    for tag in ("l_mid", "c_last"):
        case = FC.grad_case(gold, tag)
        mm = T.make_model(pkg, case["ds"])
        T.run_case(mm, case)[0].backward()
        g = T.grads_of(mm, case["lidar"])
        row.append(tag + ":" + " ".join(f"{n[:7]}={np.linalg.norm(g[n] - refs[tag][n].reshape(-1)) / max(np.linalg.norm(refs[tag][n]), 1e-30):.4f}"
                                         for n in FC.GRAD_NAMES if refs[tag][n].any()))
    print(" | ".join(row), flush=True)
