"""Where the joint training step's time goes between kernels: a torch.profiler (CUPTI) timeline of three
steps of bench.build_train_step, reduced to (a) GPU-busy time vs. step time, (b) the idle gaps and the
kernels that precede the largest ones, (c) the host time needed to enqueue one step, (d) the same step
captured in a CUDA graph and replayed (launch gaps collapse; the kernels are the same).

    python tools/prof_train_timeline.py [--rays 4096] [--graph 1]
"""
import argparse
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--profile", type=int, default=1)
a = ap.parse_args()

pkg = importlib.import_module("selfsupervised-nvsf_b200")
S = importlib.import_module("selfsupervised-nvsf_b200.synth")
dev = torch.device("cuda", 0)
cfg_kw = bench.field_kwargs(S) if hasattr(bench, "field_kwargs") else dict(
    time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND, min_near=S.MIN_NEAR,
    min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)
step, model, opt, loss_h = bench.build_train_step(pkg, S, cfg_kw, dev, 0, 1, a.rays)
for _ in range(3):
    step()
torch.cuda.synchronize()


def timed(fn, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    t_host = (time.perf_counter() - t0) / n * 1e3
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_host


ms, host_ms = timed(step)
out = {"rays": a.rays, "eager_ms_per_step": ms, "host_enqueue_ms_per_step": host_ms}

if a.profile:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if evs:
        t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
        busy, cur_end, gaps = 0.0, None, []
        prev = None
        for e in evs:
            s, en = e.time_range.start, e.time_range.end
            if cur_end is None or s > cur_end:
                if cur_end is not None:
                    gaps.append((s - cur_end, prev.name[:60], e.name[:60]))
                busy += en - s
                cur_end = en
            elif en > cur_end:
                busy += en - cur_end
                cur_end = en
            prev = e
        span = t1 - t0
        by_pair = {}
        for g, p, n in gaps:
            k = p + " -> " + n
            by_pair[k] = (by_pair.get(k, (0, 0))[0] + g, by_pair.get(k, (0, 0))[1] + 1)
        top = sorted(by_pair.items(), key=lambda kv: -kv[1][0])[:25]
        hist = {"<2us": 0, "2-5us": 0, "5-20us": 0, "20-100us": 0, ">100us": 0}
        tot = {k: 0.0 for k in hist}
        for g, _, _ in gaps:
            k = "<2us" if g < 2 else "2-5us" if g < 5 else "5-20us" if g < 20 else "20-100us" if g < 100 else ">100us"
            hist[k] += 1
            tot[k] += g
        per = {}
        for e in evs:
            k = e.name[:70]
            d = e.time_range.end - e.time_range.start
            per[k] = (per.get(k, (0.0, 0))[0] + d, per.get(k, (0.0, 0))[1] + 1)
        out["kernels_ms_per_step"] = [(k, round(v[0] / 3e3, 4), v[1] // 3)
                                      for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:40]]
        out["profile"] = {"steps": 3, "span_ms": span / 1e3, "gpu_busy_ms": busy / 1e3, "idle_ms": (span - busy) / 1e3,
                          "kernels": len(evs), "gap_count": hist, "gap_total_us": tot,
                          "top_gaps_us": [(k, round(v[0], 1), v[1]) for k, v in top]}

if a.graph:
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            step()
        torch.cuda.synchronize()
        gms, ghost = timed(g.replay)
        out["graph_ms_per_step"] = gms
        out["graph_host_ms_per_step"] = ghost
        out["graph_loss"] = float(loss_h.item())
    except Exception as e:  # noqa: BLE001
        out["graph_error"] = repr(e)[:400]
print(json.dumps(out))
