"""Diagnostic: relative L2 error of every CUDA parameter gradient vs the reference goldens."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
import field_cases as FC
import test_field_grad_gpu as T
pkg = importlib.import_module("selfsupervised-nvsf_b200")
gold = np.load(os.path.join(ROOT, "tests", "golden", "field_grad_ref.npz"))
for tag in FC.GRAD_CASES:
    case = FC.grad_case(gold, tag)
    m = T.make_model(pkg, case["ds"])
    loss, out = T.run_case(m, case)
    loss.backward()
    g = T.grads_of(m, case["lidar"])
    row = [f"{tag:8s} loss {loss.item():+.5f} ref {float(gold[tag + '_loss']):+.5f}"]
    for name in FC.GRAD_NAMES:
        k = f"{tag}_g_{name}_"
        idx, val, l2 = gold[k + "idx"], gold[k + "val"].astype(np.float64), float(gold[k + "l2"])
        if l2 == 0:
            row.append(f"{name}=zero({int(np.count_nonzero(g[name]))})")
            continue
        got = g[name].astype(np.float64)
        e1 = np.linalg.norm(got[idx] - val) / np.linalg.norm(val)
        e2 = abs(np.linalg.norm(got) - l2) / l2
        row.append(f"{name}={e1:.4f}/{e2:.4f}")
    print(" ".join(row))
