"""Host-side logic of the multi-GPU path on CPU: ray sharding and the flat-buffer gradient
all-reduce (GradSync) with the gloo backend, world_size 2."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

D = importlib.import_module("selfsupervised-nvsf_b200.dist")

SMALL = dict(device="cpu", log2_hashmap_size=8, hash_size_dynamic=(6, 5, 5), flow_log2_hashmap_size=8,
             base_resolution=16, max_resolution=64, min_resolution=4, time_resolution=2, num_frames=4,
             flow_base_resolution=4, flow_max_resolution=32)


def test_shard_range_partitions_every_count():
    for n in (0, 1, 7, 8, 4096, 67980, 529408):
        for world in (1, 2, 3, 4, 8):
            cuts = [D.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_rays_shapes():
    o = torch.arange(30, dtype=torch.float32).view(1, 10, 3)
    a, b = D.shard_rays(o, o, 1, 4)
    assert a.shape == (1, 3, 3) and torch.equal(a, o[:, 3:6])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("selfsupervised-nvsf_b200")
    torch.manual_seed(0)
    m = pkg.NeRFNetwork(**SMALL)
    sync = D.GradSync(m)
    # the parameter .grad tensors are views of one flat buffer, grouped lidar | camera | shared
    n_par = sum(p.numel() for p in m.parameters())
    assert n_par <= sync.flat.numel() < n_par + 3 * 4 * world   # groups padded to 4 * world floats
    for p in m.parameters():
        assert p.grad.untyped_storage().data_ptr() == sync.flat.untyped_storage().data_ptr()
    g = torch.Generator().manual_seed(100 + rank)
    local = torch.randn(sync.flat.numel(), generator=g)
    sync.flat.copy_(local)
    sync.reduce_group("lidar")          # e.g. while the camera render is still running
    sync.reduce_group("camera")
    sync.reduce_group("shared")
    sync.wait()
    want = sum(torch.randn(sync.flat.numel(), generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    ok = torch.allclose(sync.flat, want, atol=1e-6)
    # views still attached after zero_grad, and a bf16 wire format stays within bf16 rounding
    sync.zero_grad()
    ok = ok and all(float(p.grad.abs().sum()) == 0.0 for p in m.parameters())
    sync16 = D.GradSync(m, comm_dtype=torch.bfloat16)
    sync16.flat.copy_(local)
    sync16.reduce_all()
    ok = ok and torch.allclose(sync16.flat, want, atol=2e-2, rtol=2e-2)
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok)]))
    dist.destroy_process_group()


def test_gradsync_gloo_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert int(np.load(tmp_path / f"ok{r}.npy")[0]) == 1


def test_cell_slice_partitions_the_grid():
    for n in (8, 64, 128 ** 3, 2 * 128 ** 3, 1000 * 8):
        for world in (1, 2, 3, 4, 8):
            cuts = [D.cell_slice(n, r, world) for r in range(world)]
            per = cuts[0][0]
            assert per % 8 == 0 and per * world >= n and all(c[0] == per for c in cuts)
            assert cuts[0][1] == 0 and max(c[2] for c in cuts) == n
            covered = np.zeros(n, dtype=np.int32)
            for _, lo, hi in cuts:
                assert 0 <= lo <= hi <= n and hi - lo <= per
                covered[lo:hi] += 1
            assert (covered == 1).all()


def _torch_adam(opt, lr, grad_scale, found):
    """FlatAdam step on CPU tensors (torch.optim.Adam's update, main_nvsf.py:350-352) for the tests of the
    sharding / collective choreography; the product path is csrc/optim.cu."""
    if found is not None and float(found) != 0.0:
        opt.state[3] = 1.0
        return
    opt.state[0] += 1
    t = float(opt.state[0])
    b1, b2 = opt.betas
    for off, moff, k, scale in opt.launches:
        g = opt.sync.flat[off:off + k] * grad_scale
        m, v = opt.exp_avg[moff:moff + k], opt.exp_avg_sq[moff:moff + k]
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        opt.flat[off:off + k] -= (lr * scale / (1 - b1 ** t)) * m / (v.sqrt() / (1 - b2 ** t) ** 0.5 + opt.eps)


def _worker_sharded(rank, world, port, out_dir):
    """Sharded Adam (reduce-scatter -> Adam on the rank's slice -> all-gather) == plain Adam on the averaged
    gradient; a non-finite gradient on ONE rank makes EVERY rank skip the step."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("selfsupervised-nvsf_b200")
    O = importlib.import_module("selfsupervised-nvsf_b200.optim")
    torch.manual_seed(0)
    m = pkg.NeRFNetwork(**SMALL)
    torch.manual_seed(0)
    ref = pkg.NeRFNetwork(**SMALL)
    opt = O.FlatAdam(m, lr=1e-2, shard=True, skip_nonfinite=True, step_fn=_torch_adam)
    ropt = torch.optim.Adam(ref.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    ok = opt.sync.mode == "reduce_scatter" and opt.exp_avg.numel() * world == opt.sync.flat.numel()
    names = [n for n, _ in m.named_parameters()]
    for it in range(3):
        opt.zero_grad()
        grads = {}
        for r in range(world):
            g = torch.Generator().manual_seed(1000 * it + r)
            grads[r] = {n: torch.randn(getattr(m, n).shape, generator=g) for n in names}
        for n in names:
            getattr(m, n).grad.copy_(grads[rank][n])
        if it == 1 and rank == world - 1:     # overflow on one rank only: everybody must skip
            m.flow_mlp.grad.view(-1)[3] = float("inf")
        for grp in D.GROUPS:
            opt.sync.reduce_group(grp)
        opt.sync.wait()
        opt.step()
        if it != 1:
            for n in names:
                getattr(ref, n).grad = sum(grads[r][n] for r in range(world)) / world
            ropt.step()
    ok = ok and opt.applied_steps() == 2
    for n in names:
        ok = ok and torch.allclose(getattr(m, n).detach(), getattr(ref, n).detach(), atol=1e-6, rtol=1e-5)
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok)]))
    dist.destroy_process_group()


def test_sharded_adam_and_found_inf_gloo_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker_sharded, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert int(np.load(tmp_path / f"ok{r}.npy")[0]) == 1


def test_gradsync_single_process_is_a_noop():
    pkg = importlib.import_module("selfsupervised-nvsf_b200")
    m = pkg.NeRFNetwork(**SMALL)
    sync = D.GradSync(m)
    sync.flat.fill_(2.0)
    sync.reduce_all()
    assert float(sync.flat.min()) == 2.0 and m.fused_grad_accumulation
