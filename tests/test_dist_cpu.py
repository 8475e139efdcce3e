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
    assert sum(p.numel() for p in m.parameters()) == sync.flat.numel()
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


def test_gradsync_single_process_is_a_noop():
    pkg = importlib.import_module("selfsupervised-nvsf_b200")
    m = pkg.NeRFNetwork(**SMALL)
    sync = D.GradSync(m)
    sync.flat.fill_(2.0)
    sync.reduce_all()
    assert float(sync.flat.min()) == 2.0 and m.fused_grad_accumulation
