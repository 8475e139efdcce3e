"""Adam over the flat parameter / gradient buffers of a NeRFNetwork replica.

Mirrors the optimizer the reference builds (reference nvsf/scripts/main_nvsf.py:350-352 with the
parameter groups of NeRFNetwork.get_params, network_dynamic.py:335-357: encoders, sigma_net and
color_net at `lr`; flow_net, intensity_net and raydrop_net at 0.1 * lr), but as ONE streaming CUDA
pass per learning-rate segment (csrc/optim.cu) over buffers that already are flat: `GradSync`
(dist.py) owns the gradients, this class re-homes the parameters and the two moment buffers in the
same order.  There is no PyTorch fallback.

shard=True ("Adam fused with the all-reduce epilogue"): GradSync runs in
reduce_scatter mode, every rank keeps the Adam moments of — and updates — only its 1/world slice of each
group, and the updated parameter slices are all-gathered in place: the optimizer is the epilogue of the
gradient reduction, Adam's 28 B/parameter of HBM traffic and the moment memory shrink by the world size.

skip_nonfinite=True gives `scaler.step(optimizer)` semantics (trainer.py:1332-1334) without a host sync:
non-finite gradients are detected on the device, the 4-byte flag is all-reduced (MAX) so that every rank
agrees, and the guarded kernel leaves parameters, moments and the step count untouched."""
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .dist import GROUPS, GradSync

LR_SCALE = {"flow_grid": 0.1, "flow_mlp": 0.1, "intensity_net": 0.1, "raydrop_net": 0.1}


class FlatAdam:
    def __init__(self, model, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, grad_sync=None, shard=False,
                 skip_nonfinite=False, process_group=None, comm_dtype=None, step_fn=None):
        self.model, self.lr, self.betas, self.eps = model, float(lr), betas, float(eps)
        self.shard, self.skip_nonfinite = bool(shard), bool(skip_nonfinite)
        if grad_sync is None:
            grad_sync = GradSync(model, process_group=process_group, comm_dtype=comm_dtype,
                                 mode="reduce_scatter" if shard else "allreduce")
        if self.shard != (grad_sync.mode == "reduce_scatter"):
            raise ValueError("FlatAdam(shard=True) needs GradSync(mode='reduce_scatter') and vice versa")
        self.sync = grad_sync
        self._step_fn = step_fn   # tests of the sharding logic on CPU inject a step; None = the CUDA kernels
        params = dict(model.named_parameters())
        self.names = [n for g in GROUPS.values() for n in g]
        total = self.sync.flat.numel()
        dev = self.sync.flat.device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.segments = []
        with torch.no_grad():
            for n in self.names:
                p = params[n]
                off, k = self.sync.param_offsets[n]
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)   # parameters become views of the flat buffer
                # every tensor is padded to a multiple of 4 floats by construction of the field
                if off % 4 or k % 4:
                    raise ValueError(f"{n}: segment [{off}, {off + k}) is not 16-byte aligned")
                self.segments.append((n, off, k, LR_SCALE.get(n, 1.0)))
        # the part of the flat buffer this rank updates: its slice of every group (everything when not sharded)
        self.owned = [self.sync.shard(g) for g in GROUPS]
        n_owned = sum(hi - lo for lo, hi in self.owned)
        self.exp_avg = torch.zeros(n_owned, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n_owned, dtype=torch.float32, device=dev)
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)      # see nvsf_adam_begin
        self.found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0   # steps requested on the host; the applied count is state[0] on the device
        # merge neighbours with the same learning rate: fewer, longer launches
        merged = []
        for n, o, k, s in self.segments:
            if merged and merged[-1][3] == s and merged[-1][1] + merged[-1][2] == o:
                merged[-1] = (merged[-1][0] + "+" + n, merged[-1][1], merged[-1][2] + k, s)
            else:
                merged.append((n, o, k, s))
        # ... intersected with the owned ranges: (flat offset, moment offset, length, lr scale)
        self.launches, moff = [], 0
        for lo, hi in self.owned:
            for _, o, k, s in merged:
                a, b = max(lo, o), min(hi, o + k)
                if a < b:
                    self.launches.append((a, moff + a - lo, b - a, s))
            moff += hi - lo

    def zero_grad(self):
        self.sync.zero_grad()

    def applied_steps(self):
        """Number of steps that were not skipped (host read of the device counter)."""
        return int(self.state[0].item())

    @torch.no_grad()
    def step(self, lr=None, grad_scale=1.0):
        """One Adam step (the reference's lr scheduler passes the current lr every step,
        trainer.py:1337-1338).  The gradient reduction of every group must have been started
        (GradSync.reduce_group) and waited for (GradSync.wait)."""
        self.step_count += 1
        lr = self.lr if lr is None else float(lr)
        g = self.sync.flat
        if self._step_fn is not None:   # CPU tests of the partition / collective choreography
            found = None
            if self.skip_nonfinite:
                self.found_inf.zero_()
                for lo, hi in self.owned:
                    if not bool(torch.isfinite(g[lo:hi]).all()):
                        self.found_inf.fill_(1.0)
                found = self.sync.all_reduce_flag(self.found_inf)
            self._step_fn(self, lr, float(grad_scale), found)
        else:
            L = _lib.lib()
            st = stream_ptr()
            found = None
            if self.skip_nonfinite:
                self.found_inf.zero_()
                for lo, hi in self.owned:
                    check(L.nvsf_grad_found_inf(g.data_ptr() + 4 * lo, hi - lo, ptr(self.found_inf), st),
                          "grad_found_inf")
                found = self.sync.all_reduce_flag(self.found_inf)
            check(L.nvsf_adam_begin(ptr(self.state), ptr(found), self.betas[0], self.betas[1], st), "adam_begin")
            for off, moff, k, scale in self.launches:
                check(L.nvsf_adam_step_guarded(self.flat.data_ptr() + 4 * off, g.data_ptr() + 4 * off,
                                               self.exp_avg.data_ptr() + 4 * moff,
                                               self.exp_avg_sq.data_ptr() + 4 * moff, k, lr * scale, self.betas[0],
                                               self.betas[1], self.eps, ptr(self.state), float(grad_scale), st),
                      "adam_step_guarded")
        if self.shard:
            for gname in GROUPS:
                self.sync.all_gather_group(self.flat, gname)
        # the packed tables (fp16 hash, channel-last planes, MLP images) are stale now
        self.model._packed.clear()
