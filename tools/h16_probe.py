import importlib, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import field_cases as FC
pkg = importlib.import_module("selfsupervised-nvsf_b200")
S = FC.S
L = pkg._lib.lib()
m = pkg.NeRFNetwork(time_resolution=S.TIME_RESOLUTION, num_frames=S.NUM_FRAMES, bound=S.BOUND, min_near=S.MIN_NEAR,
                    min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH)
m.load_flat_params(FC.oracle_params()); m.eval()
for lidar in (True, False):
    o, d = (S.lidar_rays if lidar else S.camera_rays)(333, seed=41)
    to, td = torch.from_numpy(o).cuda()[None], torch.from_numpy(d).cuda()[None]
    out = {}
    for tc in (0, 2, 6):
        L.nvsf_set_option(b"heads_tc", tc)
        with torch.no_grad():
            r = m.run(to, td, torch.tensor([[0.45]], device="cuda"), cal_lidar_color=lidar, num_steps=200)
        out[tc] = r["image_lidar" if lidar else "image"].float().cpu().numpy().reshape(-1, 2 if lidar else 3)
    for tc in (2, 6):
        e = np.abs(out[tc] - out[0])
        print("lidar", lidar, "mode", tc, "max abs err", e.max(), "mean", e.mean(), "ref max", np.abs(out[0]).max(), out[tc][:2], out[0][:2])
