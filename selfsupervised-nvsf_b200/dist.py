"""Data-parallel plumbing of the ray-rendering path: one process per GPU, rays sharded by index,
parameters replicated, parameter gradients all-reduced (NCCL over NVLink / NVSwitch).

The reference has only an unused DDP hook (reference nvsf/nerf/trainer.py:82-84, never
initialised).  Rendering needs no collective at all (every kernel is per ray / per sample); a
training step needs ONE sum over ranks of the parameter gradients (SURVEY.md section 8e:
2 x 29.4 M hash + 30.5 M flow grid + 2 x 2.2 M planes + MLPs = 93.6 M fp32 = 374 MB).

GradSync owns one flat fp32 gradient buffer whose slices are the `.grad` of the parameters
(`fused_grad_accumulation`: the backward kernels accumulate straight into it, there is no
per-parameter temporary and no flatten copy).  The buffer is ordered by the moment a gradient
becomes final in a joint step — LiDAR-only parameters, camera-only parameters, shared ones — so the
all-reduce of the LiDAR group runs on a side stream while the camera render is still computing.
"""
import torch
import torch.distributed as dist

GROUPS = {
    "lidar": ("hash_static_lidar", "hash_dynamic_lidar", "planes_lidar", "intensity_net", "raydrop_net"),
    "camera": ("hash_static_camera", "hash_dynamic_camera", "planes_camera", "color_net"),
    "shared": ("flow_grid", "flow_mlp", "sigma_net"),
}


def shard_range(n, rank, world):
    """Contiguous ray-id range [lo, hi) of `rank`: sizes differ by at most 1."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_interleaved(n, rank, world, block):
    """Ray ids of `rank` when blocks of `block` consecutive rays (e.g. one image row) are dealt round-robin:
    block j goes to rank j % world.  A contiguous split of a camera frame is unbalanced under occupancy
    skipping — the upper image rows miss the scene, the lower rows do not (per-rank sample counts max / mean
    1.48 / 1.62 / 1.66 at 2 / 4 / 8 ranks on the street-shell grid) — while interleaved rows see the same mix.
    Returns a sorted int64 index tensor; every ray belongs to exactly one rank."""
    n, block = int(n), max(int(block), 1)
    nblocks = (n + block - 1) // block
    mine = torch.arange(rank, nblocks, world, dtype=torch.int64)
    idx = (mine[:, None] * block + torch.arange(block, dtype=torch.int64)[None, :]).reshape(-1)
    return idx[idx < n]


def world_and_rank(process_group=None):
    """(world size, rank) of `process_group`, (1, 0) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(process_group), dist.get_rank(process_group)
    return 1, 0


def cell_slice(n, rank, world):
    """Occupancy-grid cells of `rank` for the sharded update: (per, lo, hi) with equal
    slice length `per` = ceil(n / world) rounded up to a multiple of 8 (whole bitfield bytes, and what
    all_gather_into_tensor wants), cells [lo, hi) clipped to n.  The last ranks may own fewer (or no) cells."""
    n, world = int(n), int(world)
    per = (-(-n // world) + 7) // 8 * 8
    lo = min(rank * per, n)
    return per, lo, min(lo + per, n)


def shard_rays(rays_o, rays_d, rank, world):
    """Slice [N,3] (or [1,N,3]) ray tensors to this rank's range."""
    n = rays_o.shape[-2]
    lo, hi = shard_range(n, rank, world)
    return rays_o[..., lo:hi, :], rays_d[..., lo:hi, :]


class GradSync:
    """Flat gradient buffer + grouped, overlapped gradient reduction for a NeRFNetwork replica.

    mode "allreduce":      every rank ends with the full summed gradient (one all-reduce per group).
    mode "reduce_scatter": rank r ends with the summed gradient of slice r of every group only
                           (`shard(group)`); optim.FlatAdam(shard=True) then updates that slice of the
                           parameters and all-gathers the parameters — the optimizer as the epilogue of the
                           reduction: the same bytes on the wire as a ring
                           all-reduce, Adam traffic and moment memory divided by the world size.
    Groups are padded to a multiple of 4 * world floats so that every slice is 16-byte aligned."""

    def __init__(self, model, process_group=None, average=True, comm_dtype=None, mode="allreduce"):
        if mode not in ("allreduce", "reduce_scatter"):
            raise ValueError(f"GradSync mode {mode!r}")
        self.model, self.pg, self.average, self.mode = model, process_group, average, mode
        self.comm_dtype = comm_dtype  # e.g. torch.bfloat16 halves the bytes on the wire
        self.world, self.rank = world_and_rank(process_group)
        params = dict(model.named_parameters())
        missing = [n for g in GROUPS.values() for n in g if n not in params]
        if missing:
            raise KeyError(f"model lacks parameters {missing}")
        dev = next(iter(params.values())).device
        quantum = 4 * self.world
        self.slices, self.param_offsets, off = {}, {}, 0
        for gname, names in GROUPS.items():
            start = off
            for n in names:
                self.param_offsets[n] = (off, params[n].numel())
                off += params[n].numel()
            off = start + -(-(off - start) // quantum) * quantum
            self.slices[gname] = (start, off)
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for n, (o, k) in self.param_offsets.items():
            params[n].grad = self.flat[o:o + k].view_as(params[n])
        model.fused_grad_accumulation = True
        self.cuda = dev.type == "cuda"
        self.stream = torch.cuda.Stream(device=dev) if self.cuda else None
        self._pending = []

    def zero_grad(self):
        """One memset of the flat buffer (the views stay attached to the parameters)."""
        self.flat.zero_()

    def group_view(self, group):
        a, b = self.slices[group]
        return self.flat[a:b]

    def shard(self, group, rank=None):
        """[lo, hi) of this rank's slice of `group` in the flat buffer (the whole group in all-reduce mode)."""
        a, b = self.slices[group]
        if self.mode != "reduce_scatter":
            return a, b
        per = (b - a) // self.world
        r = self.rank if rank is None else rank
        return a + r * per, a + (r + 1) * per

    def reduce_group(self, group):
        """Start the reduction of one group; its gradients must be final on the current stream."""
        if self.world == 1:
            return
        buf = self.group_view(group)
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.stream.wait_event(ev)
            ctx = torch.cuda.stream(self.stream)
        else:
            ctx = _null()
        with ctx:
            wire = buf if self.comm_dtype in (None, torch.float32) else buf.to(self.comm_dtype)
            if self.mode == "reduce_scatter":
                lo, hi = self.shard(group)
                mine = self.flat[lo:hi]
                # NCCL reduces in place when the output is the rank's own slice of the input
                inplace = wire is buf and self.cuda and dist.get_backend(self.pg) == "nccl"
                out = mine if inplace else torch.empty(hi - lo, dtype=wire.dtype, device=wire.device)
                dist.reduce_scatter_tensor(out, wire, op=dist.ReduceOp.SUM, group=self.pg)
                if not inplace:
                    mine.copy_(out)
                if self.average:
                    mine.mul_(1.0 / self.world)
            else:
                dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.pg)
                if wire is not buf:
                    buf.copy_(wire)
                if self.average:
                    buf.mul_(1.0 / self.world)
        self._pending.append(group)

    def wait(self):
        """Make the current stream wait for every started reduction."""
        if self.cuda and self._pending:
            torch.cuda.current_stream().wait_stream(self.stream)
        self._pending.clear()

    def reduce_all(self):
        for g in GROUPS:
            self.reduce_group(g)
        self.wait()

    def all_reduce_flag(self, flag):
        """MAX over ranks of a 4-byte device flag (found_inf of the AMP step, trainer.py:1332-1334): every
        replica must take the same skip / apply decision or the parameters diverge."""
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.pg)
        return flag

    def all_gather_group(self, flat_params, group):
        """In-place all-gather of a parameter buffer laid out like `flat`: every rank contributes its
        slice of `group` and receives the others'."""
        if self.world == 1 or self.mode != "reduce_scatter":
            return
        a, b = self.slices[group]
        lo, hi = self.shard(group)
        dist.all_gather_into_tensor(flat_params[a:b], flat_params[lo:hi].clone() if not self.cuda
                                    else flat_params[lo:hi], group=self.pg)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
