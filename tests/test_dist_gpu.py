"""2-GPU NCCL parity of the data-parallel training path (BASELINE configs[4]): the gradients
every rank holds after GradSync's grouped all-reduce equal the single-GPU gradients of the concatenated
batch (the reference's only multi-GPU precedent is DDP's gradient average, trainer.py:82-84).
Needs two visible GPUs (`gpurun --gpus 2`); skipped otherwise."""
import importlib
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

import field_cases as FC

S = FC.S
N_PER_RANK, STEPS, T = 512, 128, 0.4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _render_loss(pkg, m, o, d, noise, lidar, dev):
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = m.render(to(o)[None], to(d)[None], torch.tensor([[T]], device=dev), cal_lidar_color=lidar, staged=False,
                   num_steps=STEPS, noise=to(noise))
    sfx = "_lidar" if lidar else ""
    return out["depth" + sfx].sum() + out["image" + sfx].sum() + 0.5 * out["weights_sum" + sfx].sum()


def _batch(world):
    lo, ld = S.lidar_rays(N_PER_RANK * world, seed=31)
    co, cd = S.camera_rays(N_PER_RANK * world, seed=32)
    rng = np.random.default_rng(33)
    return lo, ld, co, cd, rng.random((N_PER_RANK * world, STEPS), dtype=np.float32)


def _make(pkg, dev):
    m = pkg.NeRFNetwork(device=dev, time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                        min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH).train()
    m.load_flat_params(FC.oracle_params())
    return m


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pkg = importlib.import_module("selfsupervised-nvsf_b200")
    D = pkg.dist
    lo, ld, co, cd, noise = _batch(world)
    a, b = D.shard_range(N_PER_RANK * world, rank, world)
    m = _make(pkg, dev)
    sync = D.GradSync(m, average=False)            # sum over ranks == gradient of the concatenated batch
    _render_loss(pkg, m, lo[a:b], ld[a:b], noise[a:b], True, dev).backward()
    sync.reduce_group("lidar")                     # overlaps the camera render, as in the training step
    _render_loss(pkg, m, co[a:b], cd[a:b], noise[a:b], False, dev).backward()
    sync.reduce_group("camera")
    sync.reduce_group("shared")
    sync.wait()
    torch.cuda.synchronize()
    got = sync.flat.detach().cpu().numpy().astype(np.float64)
    # every rank must hold the same bits
    chk = torch.tensor([float(np.abs(got).sum())], device=dev, dtype=torch.float64)
    lo_, hi_ = chk.clone(), chk.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    same = float(lo_) == float(hi_)
    res = {"same": same}
    if rank == 0:
        # single GPU, concatenated batch
        m1 = _make(pkg, dev)
        s1 = D.GradSync(m1, average=False)
        _render_loss(pkg, m1, lo, ld, noise, True, dev).backward()
        _render_loss(pkg, m1, co, cd, noise, False, dev).backward()
        torch.cuda.synchronize()
        want = s1.flat.detach().cpu().numpy().astype(np.float64)
        for g, (x, y) in s1.slices.items():
            res[g] = float(np.linalg.norm(got[x:y] - want[x:y]) / np.linalg.norm(want[x:y]))
    np.save(os.path.join(out_dir, f"res{rank}.npy"), np.array([res], dtype=object), allow_pickle=True)
    dist.destroy_process_group()


def test_allreduced_gradients_equal_single_gpu_gradients(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "res0.npy", allow_pickle=True)[0]
    r1 = np.load(tmp_path / "res1.npy", allow_pickle=True)[0]
    assert r0["same"] and r1["same"]
    print("rel-L2 of all-reduced vs single-GPU gradients per group:", {k: v for k, v in r0.items() if k != "same"})
    # same kernels, same samples; only the order of fp32 atomic accumulation and of the cross-rank sum differs
    for g in ("lidar", "camera", "shared"):
        assert r0[g] < 1e-4, (g, r0[g])


def _worker_sharded(rank, world, port, out_dir):
    """(1) reduce-scatter -> Adam on the rank's slice -> all-gather (FlatAdam(shard=True), bf16 and fp32 wire)
    against all-reduce -> full Adam; one rank's overflow makes every rank skip.  (2) occupancy-grid update with
    the cells sharded over the ranks against the single-GPU update with the same jitter."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pkg = importlib.import_module("selfsupervised-nvsf_b200")
    res = {}
    small = dict(device=dev, log2_hashmap_size=12, hash_size_dynamic=(10, 9, 9), flow_log2_hashmap_size=11,
                 time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND, min_near=S.MIN_NEAR,
                 min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)
    models, opts = [], []
    for shard in (False, True):
        torch.manual_seed(0)
        m = pkg.NeRFNetwork(**small).train()
        models.append(m)
        opts.append(pkg.optim.FlatAdam(m, lr=1e-2, shard=shard, skip_nonfinite=True))
    names = [n for n, _ in models[0].named_parameters()]
    for it in range(4):
        g = torch.Generator(device=dev).manual_seed(10 * it + rank)
        grads = {n: torch.randn(getattr(models[0], n).shape, generator=g, device=dev) for n in names}
        if it == 2 and rank == world - 1:
            grads["planes_camera"].view(-1)[7] = float("inf")
        for m, opt in zip(models, opts):
            opt.zero_grad()
            for n in names:
                getattr(m, n).grad.copy_(grads[n])
            for grp in pkg.dist.GROUPS:
                opt.sync.reduce_group(grp)
            opt.sync.wait()
            opt.step()
    torch.cuda.synchronize()
    res["applied"] = (opts[0].applied_steps(), opts[1].applied_steps())
    res["moment_fraction"] = opts[1].exp_avg.numel() / opts[1].sync.flat.numel()
    res["adam"] = max(float((getattr(models[0], n).detach() - getattr(models[1], n).detach()).abs().max())
                      for n in names)
    # every rank holds the same parameters after the all-gather
    chk = torch.stack([getattr(models[1], n).detach().double().sum() for n in names])
    lo_, hi_ = chk.clone(), chk.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    res["same_params"] = bool(torch.equal(lo_, hi_))

    # (2) sharded occupancy-grid update
    m = _make(pkg, dev).eval()
    n = m.cascade * m.grid_size ** 3
    noise = torch.rand(n, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    bits_sharded = m.update_extra_state([0.1, 0.6], cal_lidar_color=True, noise=noise).clone()
    grid_sharded = m.density_grid(True).clone()
    m1 = _make(pkg, dev).eval()
    bits_single = m1.update_extra_state([0.1, 0.6], cal_lidar_color=True, noise=noise, shard=False)
    torch.cuda.synchronize()
    res["grid_equal"] = bool(torch.equal(grid_sharded, m1.density_grid(True)))
    res["bits_equal"] = bool(torch.equal(bits_sharded, bits_single))
    res["occupied"] = int(torch.count_nonzero(bits_single))
    np.save(os.path.join(out_dir, f"res{rank}.npy"), np.array([res], dtype=object), allow_pickle=True)
    dist.destroy_process_group()


def test_sharded_adam_and_sharded_grid_update(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    mp.spawn(_worker_sharded, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = np.load(tmp_path / f"res{r}.npy", allow_pickle=True)[0]
        print(res)
        assert res["applied"] == (3, 3)                    # step 2 skipped on BOTH ranks, both optimizers
        assert abs(res["moment_fraction"] - 1.0 / world) < 1e-6
        assert res["adam"] < 1e-6 and res["same_params"]   # differs only by the summation order of the reduction
        assert res["grid_equal"] and res["bits_equal"] and res["occupied"] > 0
