"""Adam over the flat parameter / gradient buffers of a NeRFNetwork replica.

Mirrors the optimizer the reference builds (reference nvsf/scripts/main_nvsf.py:350-352 with the
parameter groups of NeRFNetwork.get_params, network_dynamic.py:335-357: encoders, sigma_net and
color_net at `lr`; flow_net, intensity_net and raydrop_net at 0.1 * lr), but as ONE streaming CUDA
pass per learning-rate segment (csrc/optim.cu) over buffers that already are flat: `GradSync`
(dist.py) owns the gradients, this class re-homes the parameters and the two moment buffers in the
same order.  There is no PyTorch fallback."""
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .dist import GROUPS, GradSync

LR_SCALE = {"flow_grid": 0.1, "flow_mlp": 0.1, "intensity_net": 0.1, "raydrop_net": 0.1}


class FlatAdam:
    def __init__(self, model, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, grad_sync=None):
        self.model, self.lr, self.betas, self.eps = model, float(lr), betas, float(eps)
        self.sync = grad_sync if grad_sync is not None else GradSync(model)
        params = dict(model.named_parameters())
        self.names = [n for g in GROUPS.values() for n in g]
        total = self.sync.flat.numel()
        dev = self.sync.flat.device
        self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.segments, off = [], 0
        with torch.no_grad():
            for n in self.names:
                p = params[n]
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)   # parameters become views of the flat buffer
                # every tensor is padded to a multiple of 4 floats by construction of the field
                if off % 4 or k % 4:
                    raise ValueError(f"{n}: segment [{off}, {off + k}) is not 16-byte aligned")
                self.segments.append((n, off, k, LR_SCALE.get(n, 1.0)))
                off += k
        self.step_count = 0
        # merge neighbours with the same learning rate: fewer, longer launches
        merged = []
        for n, o, k, s in self.segments:
            if merged and merged[-1][3] == s and merged[-1][1] + merged[-1][2] == o:
                merged[-1] = (merged[-1][0] + "+" + n, merged[-1][1], merged[-1][2] + k, s)
            else:
                merged.append((n, o, k, s))
        self.launches = merged

    def zero_grad(self):
        self.sync.zero_grad()

    @torch.no_grad()
    def step(self, lr=None, grad_scale=1.0):
        """One Adam step (the reference's lr scheduler passes the current lr every step,
        trainer.py:1337-1338)."""
        L = _lib.lib()
        self.step_count += 1
        lr = self.lr if lr is None else float(lr)
        g = self.sync.flat
        st = stream_ptr()
        for _, off, k, scale in self.launches:
            check(L.nvsf_adam_step(self.flat.data_ptr() + 4 * off, g.data_ptr() + 4 * off,
                                   self.exp_avg.data_ptr() + 4 * off, self.exp_avg_sq.data_ptr() + 4 * off, k,
                                   lr * scale, self.betas[0], self.betas[1], self.eps, self.step_count,
                                   float(grad_scale), st), "adam_step")
        # the packed tables (fp16 hash, channel-last planes, MLP images) are stale now
        self.model._packed.clear()
