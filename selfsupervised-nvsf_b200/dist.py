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


def shard_rays(rays_o, rays_d, rank, world):
    """Slice [N,3] (or [1,N,3]) ray tensors to this rank's range."""
    n = rays_o.shape[-2]
    lo, hi = shard_range(n, rank, world)
    return rays_o[..., lo:hi, :], rays_d[..., lo:hi, :]


class GradSync:
    """Flat gradient buffer + grouped, overlapped all-reduce for a NeRFNetwork replica."""

    def __init__(self, model, process_group=None, average=True, comm_dtype=None):
        self.model, self.pg, self.average = model, process_group, average
        self.comm_dtype = comm_dtype  # e.g. torch.bfloat16 halves the bytes on the wire
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = dict(model.named_parameters())
        missing = [n for g in GROUPS.values() for n in g if n not in params]
        if missing:
            raise KeyError(f"model lacks parameters {missing}")
        dev = next(iter(params.values())).device
        total = sum(params[n].numel() for g in GROUPS.values() for n in g)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.slices, off = {}, 0
        for gname, names in GROUPS.items():
            start = off
            for n in names:
                p = params[n]
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.slices[gname] = (start, off)
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

    def reduce_group(self, group):
        """Start the all-reduce of one group; its gradients must be final on the current stream."""
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
            if self.comm_dtype is not None and self.comm_dtype != torch.float32:
                wire = buf.to(self.comm_dtype)
                dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.pg)
                buf.copy_(wire)
            else:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)
            if self.average:
                buf.mul_(1.0 / self.world)
        self._pending.append(group)

    def wait(self):
        """Make the current stream wait for every started all-reduce."""
        if self.cuda and self._pending:
            torch.cuda.current_stream().wait_stream(self.stream)
        self._pending.clear()

    def reduce_all(self):
        for g in GROUPS:
            self.reduce_group(g)
        self.wait()


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
