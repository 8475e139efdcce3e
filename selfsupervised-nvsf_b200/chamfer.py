"""Chamfer distance — drop-in for the reference module `nvsf/nerf/chamfer3D/dist_chamfer_3D.py`
(`chamfer_3DFunction`, `chamfer_3DDist`): same call shape and results, over the C ABI
(`nvsf_chamfer_forward` / `nvsf_chamfer_backward`, csrc/chamfer.cu).  GPU tensors only, as the
reference ("GPU tensors only", dist_chamfer_3D.py:41); no CPU fallback."""
import torch
from torch import nn
from torch.autograd import Function

from ._lib import NvsfError, check, lib, ptr, stream_ptr


class chamfer_3DFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        if not (xyz1.is_cuda and xyz2.is_cuda):
            raise NvsfError("chamfer_3DDist: GPU tensors only")
        if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.size(2) != 3 or xyz2.size(2) != 3:
            raise AssertionError("Wrong last dimension for the chamfer distance 's input! Check with .size()")
        a = xyz1.detach().to(dtype=torch.float32).contiguous()
        b = xyz2.detach().to(dtype=torch.float32).contiguous()
        B, n, m = a.size(0), a.size(1), b.size(1)
        dev = a.device
        # every element is written by the kernels (the reference wrapper zero-fills, dist_chamfer_3D.py:50-55)
        dist1 = torch.empty(B, n, dtype=torch.float32, device=dev)
        dist2 = torch.empty(B, m, dtype=torch.float32, device=dev)
        idx1 = torch.empty(B, n, dtype=torch.int32, device=dev)
        idx2 = torch.empty(B, m, dtype=torch.int32, device=dev)
        L = lib()
        wb = L.nvsf_chamfer_workspace_bytes(B, n, m)
        ws = torch.empty(max(wb, 8), dtype=torch.uint8, device=dev)
        check(L.nvsf_chamfer_forward(ptr(a), ptr(b), B, n, m, ptr(dist1), ptr(dist2), ptr(idx1), ptr(idx2), ptr(ws),
                                     wb, stream_ptr()), "chamfer_forward")
        ctx.save_for_backward(a, b, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        a, b, idx1, idx2 = ctx.saved_tensors
        g1 = graddist1.to(dtype=torch.float32).contiguous()
        g2 = graddist2.to(dtype=torch.float32).contiguous()
        gradxyz1, gradxyz2 = torch.zeros_like(a), torch.zeros_like(b)
        check(lib().nvsf_chamfer_backward(ptr(a), ptr(b), a.size(0), a.size(1), b.size(1), ptr(g1), ptr(g2),
                                          ptr(idx1), ptr(idx2), ptr(gradxyz1), ptr(gradxyz2), stream_ptr()),
              "chamfer_backward")
        return gradxyz1, gradxyz2


class chamfer_3DDist(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, input1, input2):
        input1 = input1.contiguous()
        input2 = input2.contiguous()
        return chamfer_3DFunction.apply(input1, input2)
