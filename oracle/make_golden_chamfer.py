"""TEST INFRASTRUCTURE.  Generates tests/golden/chamfer_ref_sm100a.npz by running the reference's
OWN Chamfer extension (oracle/_ref/chamfer_3D_ref.so = unmodified nvsf/nerf/chamfer3D/chamfer3D.cu +
chamfer_cuda.cpp built for sm_100a by oracle/build_ref_chamfer.sh) on seeded inputs.  Needs a GPU:
    gpurun -- python oracle/make_golden_chamfer.py gpurun_out/golden
then copy the .npz into tests/golden/.  tests/test_chamfer_oracle_golden.py pins the C oracle
against these vectors on CPU; tests/test_chamfer_gpu.py pins the CUDA product against them."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import chamfer_cases as CC  # noqa: E402


def load_ref():
    path = os.path.join(ROOT, "oracle", "_ref", "chamfer_3D_ref.so")
    spec = importlib.util.spec_from_file_location("chamfer_3D_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_ref(ref, a, b, g1, g2):
    """dist_chamfer_3D.py:44-88, on the reference extension."""
    xa, xb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    B, n, m = xa.shape[0], xa.shape[1], xb.shape[1]
    d1, d2 = torch.zeros(B, n).cuda(), torch.zeros(B, m).cuda()
    i1, i2 = torch.zeros(B, n).type(torch.IntTensor).cuda(), torch.zeros(B, m).type(torch.IntTensor).cuda()
    ref.forward(xa, xb, d1, d2, i1, i2)
    ga, gb = torch.zeros(xa.size()).cuda(), torch.zeros(xb.size()).cuda()
    ref.backward(xa, xb, ga, gb, torch.from_numpy(g1).cuda(), torch.from_numpy(g2).cuda(), i1, i2)
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in (d1, d2, i1, i2, ga, gb)]


def main(out_dir):
    ref = load_ref()
    out = {}
    for name in CC.GOLDEN_CASES:
        a, b, g1, g2 = CC.case(name)
        d1, d2, i1, i2, ga, gb = run_ref(ref, a, b, g1, g2)
        for k, v in zip(("dist1", "dist2", "idx1", "idx2", "grad1", "grad2"), (d1, d2, i1, i2, ga, gb)):
            out[f"{name}_{k}"] = v
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, "chamfer_ref_sm100a.npz"), **out)
    print("wrote", os.path.join(out_dir, "chamfer_ref_sm100a.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
