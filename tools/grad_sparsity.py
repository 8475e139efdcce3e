import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
import field_cases as FC
import test_field_grad_gpu as T
pkg = importlib.import_module("selfsupervised-nvsf_b200")
gold = np.load(os.path.join(ROOT, "tests", "golden", "field_grad_ref.npz"))
case = FC.grad_case(gold, "l_mid")
m = T.make_model(pkg, case["ds"])
loss, _ = T.run_case(m, case); loss.backward()
g = T.grads_of(m, True)
e, _, _ = FC.oracle_grads(case)
for name in ("hash_static", "hash_dynamic", "flow_grid"):
    ref = e[name].reshape(-1); got = g[name]
    extra = (ref == 0) & (got != 0); miss = (ref != 0) & (got == 0)
    print(name, "nnz ref", int((ref != 0).sum()), "ours", int((got != 0).sum()), "extra", int(extra.sum()), "missing", int(miss.sum()),
          "norm extra/ref", float(np.linalg.norm(got[extra]) / np.linalg.norm(ref)), "max extra", float(np.abs(got[extra]).max() if extra.any() else 0),
          "max ref", float(np.abs(ref).max()))
    idx = np.flatnonzero(extra)[:12]
    print("   extra idx", idx, got[idx])
