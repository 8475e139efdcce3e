"""Experiment: one rank's 1/8 shard of the camera frame (occupancy skipping) — eager run_cuda vs the same frame
captured in a CUDA graph and replayed: how much of the 2.4 ms per frame at 8 ranks is launch / host time?"""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("selfsupervised-nvsf_b200")
S = importlib.import_module("selfsupervised-nvsf_b200.synth")
dev = torch.device("cuda", 0)
model = pkg.NeRFNetwork(device=dev, time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND,
                        min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH).eval()
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
o, d = S.camera_rays(-1, seed=0)
idx = pkg.dist.shard_interleaved(o.shape[0], 0, world, S.CAM_W).numpy()
o_pin = torch.from_numpy(np.ascontiguousarray(o[idx]))[None].pin_memory()
d_pin = torch.from_numpy(np.ascontiguousarray(d[idx]))[None].pin_memory()
img_h = torch.empty(1, idx.shape[0], 3).pin_memory()
bits = torch.from_numpy(S.packbits_np(S.density_grid("shell"), 0.01)).to(dev)
t = torch.tensor([[0.5]], device=dev)
model.run_cuda(o_pin.to(dev), d_pin.to(dev), t, cal_lidar_color=False, dt_gamma=S.DT_GAMMA, T_thresh=1e-2,
               density_bitfield=bits, one_shot=True)
cap = (int(model.last_run_cuda_samples * 1.02) + 127) // 128 * 128
ro_s, rd_s = torch.empty_like(o_pin, device=dev), torch.empty_like(d_pin, device=dev)


def step():
    ro_s.copy_(o_pin, non_blocking=True); rd_s.copy_(d_pin, non_blocking=True)
    r = model.run_cuda(ro_s, rd_s, t, cal_lidar_color=False, dt_gamma=S.DT_GAMMA, T_thresh=1e-2,
                       density_bitfield=bits, one_shot=True, sample_capacity=cap)
    img_h.copy_(r["image"], non_blocking=True)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    host = (time.perf_counter() - t0) / n * 1e3
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host


out = {"world": world, "rays": int(idx.shape[0]), "samples": int(model.last_run_cuda_samples)}
out["eager_ms"], out["eager_host_ms"] = timed(step)
ref = img_h.clone()
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step(); step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        step()
    torch.cuda.synchronize()
    out["graph_ms"], out["graph_host_ms"] = timed(g.replay)
    out["graph_equal"] = bool(torch.equal(ref, img_h))
except Exception as e:  # noqa: BLE001
    out["graph_error"] = repr(e)[:300]
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
per = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name[:64]
        per[k] = (per.get(k, (0.0, 0))[0] + (e.time_range.end - e.time_range.start), per.get(k, (0.0, 0))[1] + 1)
out["kernels_us_per_frame"] = [(k, round(v[0] / 5, 1), v[1] // 5) for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:16]]
print(json.dumps(out))
