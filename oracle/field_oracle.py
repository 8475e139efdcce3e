"""TEST INFRASTRUCTURE.  CPU oracle ("port") of the reference's per-sample field and of its
uniform-sample renderer, in plain PyTorch fp32 on flat parameter tensors.

Restates (does not import) the reference:
  NeRFNetwork.density   nvsf/nerf/models/network_dynamic.py:213-287
  NeRFNetwork.color     network_dynamic.py:290-332
  NeRFNetwork.flow      network_dynamic.py:197-211
  HashGridT / HashGrid4D  nvsf/nerf/models/hash_field.py:65-88,143-173
  Planes4D              nvsf/nerf/models/planes_field.py:55-140
  FlowField             nvsf/nerf/models/flow_field.py:105-133
  trunc_exp             nvsf/nerf/activation.py:6-20
  NeRFRenderer.run/render  nvsf/nerf/models/renderer_dynamic.py:109-326
The tcnn pieces come from oracle/tcnn_standin.py (PARITY UNPINNED for those, see there).

Pinned by tests/test_field_oracle_golden.py against outputs of the reference's own modules
imported by file path (oracle/make_golden_field.py), on parameters produced by
oracle/field_init.py.  It is also the CPU baseline timed by bench.py.
"""
import itertools
import math

import numpy as np
import torch
import torch.nn.functional as Fnn

from . import tcnn_standin as T

PLANE_COMBS = list(itertools.combinations(range(4), 2))  # (0,1),(0,2),(0,3),(1,2),(1,3),(2,3)


class _TruncExp(torch.autograd.Function):
    """trunc_exp (activation.py:6-20): exp forward, gradient g * exp(clamp(x, -15, 15))."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(-15, 15))


class FieldConfig:
    """Sizes of the reference field at its defaults (main_nvsf.py:44-60, hash_field.py:95-101,
    flow_field.py:50-55) — every number can be shrunk for tests."""

    def __init__(self, bound=2.0, num_frames=64, time_resolution=8, min_resolution=32,
                 n_levels_plane=4, n_features_plane=8, base_resolution=512, max_resolution=32768,
                 n_levels_hash=8, n_features_hash=4, log2_hashmap_size=19,
                 hash_size_dynamic=(15, 13, 13), flow_levels=16, flow_features=8, flow_base=32,
                 flow_max=8192, flow_log2=18, hidden=64, geo_feat_dim=15, density_scale=1.0,
                 min_near=0.01, min_near_lidar=0.01, lidar_max_depth=0.81, active_sensor=False):
        self.__dict__.update(locals())
        del self.__dict__["self"]
        self.hash_pls = float(np.exp2(np.log2(max_resolution / base_resolution) / (n_levels_hash - 1)))
        self.flow_pls = float(np.exp2(np.log2(flow_max / flow_base) / (flow_levels - 1)))
        self.static_levels, self.static_entries = T.grid_levels(3, n_levels_hash, base_resolution,
                                                                self.hash_pls, log2_hashmap_size)
        self.dyn_levels, self.dyn_entries = [], []
        for h in hash_size_dynamic:
            lv, tot = T.grid_levels(2, n_levels_hash, base_resolution, self.hash_pls, h)
            self.dyn_levels.append(lv)
            self.dyn_entries.append(tot)
        self.flow_grid_levels, self.flow_entries = T.grid_levels(3, flow_levels, flow_base, self.flow_pls,
                                                                 flow_log2)
        self.plane_res = []  # per scale: (X, Y, Z, T) resolution
        for s in range(n_levels_plane):
            m = 2 ** s
            self.plane_res.append((min_resolution * m,) * 3 + (time_resolution,))

    # ---- flat parameter sizes (fp32 elements)
    def sizes(self):
        F, Tn = self.n_features_hash, self.time_resolution
        planes = sum(self.n_features_plane * r[a] * r[b] for r in self.plane_res for (a, b) in PLANE_COMBS)
        h = self.hidden
        return dict(
            hash_static=self.static_entries * F,
            hash_dynamic=sum(Tn * e * F for e in self.dyn_entries),
            planes=planes,
            flow_grid=self.flow_entries * self.flow_features,
            flow_mlp=h * (self.flow_levels * self.flow_features // 4) + h * h + 6 * h,
            sigma_net=h * 128 + 16 * h,
            intensity_net=h * 96 + h * h + 16 * h,
            raydrop_net=h * 96 + h * h + 16 * h,
            color_net=h * 32 + h * h + 16 * h,
        )


def lagrange4(t):
    """Cubic Lagrange basis on nodes {0,1/3,2/3,1} (hash_field.py:65-74, flow_field.py:105-114)."""
    nodes = [i / 3 for i in range(4)]
    return [math.prod((t - nodes[m]) / (nodes[j] - nodes[m]) for m in range(4) if m != j) for j in range(4)]


def hash_grid(table_flat, levels, D, F, x):
    """tcnn HashGrid forward (see tcnn_standin._HashGrid) on a flat fp32 table."""
    g = T._HashGrid.__new__(T._HashGrid)
    torch.nn.Module.__init__(g)
    g.D, g.L, g.F, g.levels = D, len(levels), F, levels
    g.params = table_flat
    return T._HashGrid.forward(g, x)


def mlp(flat, shapes, x, n_out, fp16_weights=True):
    """Half-precision MLP as the reference runs it: tcnn FullyFusedMLP (fp16 weights, fp16 activations
    between layers, fp16 output) or nn.Linear layers under fp16 autocast (trainer.py:1318 with
    configs/kitti360_1908.txt:23).  Storage points are rounded to fp16, the products accumulate in fp32."""
    w = T._fp16_round(flat) if fp16_weights else flat
    pad = shapes[0][1] - x.shape[1]
    h = Fnn.pad(x, (0, pad), value=1.0) if pad else x   # tcnn pads Network inputs with 1 (tcnn_standin.Network)
    h = T._fp16_round(h)
    off = 0
    for li, (o, i) in enumerate(shapes):
        W = w[off:off + o * i].view(o, i)
        off += o * i
        h = h @ W.t()
        if li != len(shapes) - 1:
            h = torch.relu(h)
        h = T._fp16_round(h)
    return h[:, :n_out]


class FieldOracle:
    """params: dict of flat fp32 torch tensors:
        {lidar,camera}: hash_static, hash_dynamic, planes ; shared: flow_grid, flow_mlp, sigma_net,
        intensity_net, raydrop_net, color_net  (layout: oracle/field_init.py)."""

    def __init__(self, cfg, params):
        self.cfg, self.p = cfg, params

    # ---------------------------------------------------------------- encoders
    def _hash_dynamic(self, mod, x, t):
        """HashGrid4D.forward_dynamic (hash_field.py:148-159) with HashGridT.forward (:76-88)."""
        c = self.cfg
        F, L, Tn = c.n_features_hash, c.n_levels_hash, c.time_resolution
        flat = self.p[mod]["hash_dynamic"]
        idx = np.float32(t) * np.float32(Tn - 1)
        k1, k2 = int(math.floor(idx)), int(math.ceil(idx))
        basis = lagrange4(float(np.float32(t)))
        outs, base = [], 0
        for pi, dims in enumerate(([0, 1], [0, 2], [1, 2])):
            per = c.dyn_entries[pi] * F
            x2 = x[:, dims]
            g1 = hash_grid(flat[base + k1 * per: base + (k1 + 1) * per], c.dyn_levels[pi], 2, F, x2)
            if k1 == k2:
                feat = g1
            else:
                g2 = hash_grid(flat[base + k2 * per: base + (k2 + 1) * per], c.dyn_levels[pi], 2, F, x2)
                feat = float(k2 - idx) * g1 + float(idx - k1) * g2
            feat = feat.view(-1, L, F)
            outs.append(sum(basis[i] * feat[:, :, i] for i in range(4)))  # [N, L]
            base += Tn * per
        return torch.cat(outs, dim=-1)  # [N, 3L]

    def _planes(self, mod, xt, which):
        """Planes4D multi-scale features (planes_field.py:86-140): product over the three space
        planes ('static') or the three time planes ('dynamic'), scales concatenated."""
        c = self.cfg
        flat = self.p[mod]["planes"]
        off, outs = 0, []
        for r in c.plane_res:
            prod = None
            for (a, b) in PLANE_COMBS:
                n = c.n_features_plane * r[a] * r[b]
                is_time = 3 in (a, b)
                if (which == "dynamic") == is_time:
                    grid = flat[off:off + n].view(1, c.n_features_plane, r[b], r[a])
                    coords = (xt[:, [a, b]] * 2.0 - 1.0).view(1, 1, -1, 2)
                    s = Fnn.grid_sample(grid, coords, mode="bilinear", padding_mode="border", align_corners=True)
                    s = s.view(c.n_features_plane, -1).t()
                    prod = s if prod is None else prod * s
                off += n
            outs.append(prod)
        return torch.cat(outs, dim=-1)

    def flow_net(self, xn, t):
        """FlowField.forward (flow_field.py:116-133): t is ONE scalar for the whole batch (:125)."""
        c = self.cfg
        enc = hash_grid(self.p["flow_grid"], c.flow_grid_levels, 3, c.flow_features, xn)
        enc = enc.view(-1, c.flow_levels, c.flow_features)
        basis = lagrange4(float(np.float32(t)))
        nb = c.flow_features // 4
        h = sum(basis[i] * enc[:, :, i * nb:(i + 1) * nb] for i in range(4)).reshape(xn.shape[0], -1)
        hdim = c.hidden
        shapes = [(hdim, h.shape[1]), (hdim, hdim), (6, hdim)]
        return mlp(self.p["flow_mlp"], shapes, h, 6)

    def flow(self, x, t):
        xn = (x + self.cfg.bound) / (2 * self.cfg.bound)
        f = self.flow_net(xn, t)
        return {"flow_forward": f[:, :3], "flow_backward": f[:, 3:]}

    # ---------------------------------------------------------------- density / color
    def features(self, x, t, lidar):
        c = self.cfg
        mod = "lidar" if lidar else "camera"
        t = float(np.float32(t))
        xn = (x + c.bound) / (2 * c.bound)
        frame_idx = int(np.float32(t) * np.float32(c.num_frames - 1))
        hash_s = hash_grid(self.p[mod]["hash_static"], c.static_levels, 3, c.n_features_hash, xn)
        hash_d = self._hash_dynamic(mod, xn, t)
        tcol = torch.full((xn.shape[0], 1), t, dtype=torch.float32)
        xt = torch.cat([xn, tcol], dim=-1)
        plane_s = self._planes(mod, xt, "static")
        plane_d = self._planes(mod, xt, "dynamic")
        flow = self.flow_net(xn, t)
        hash_1 = hash_2 = hash_d
        plane_1 = plane_2 = plane_d
        if frame_idx < c.num_frames - 1:
            x1 = xn + flow[:, :3]
            t1 = float(np.float32((frame_idx + 1) / c.num_frames))
            with torch.no_grad():  # network_dynamic.py:245-249: the warped hash query carries no gradient
                hash_1 = self._hash_dynamic(mod, x1, t1)
            plane_1 = self._planes(mod, torch.cat([x1, torch.full_like(tcol, t1)], -1), "dynamic")
        if frame_idx > 0:
            x2 = xn + flow[:, 3:]
            t2 = float(np.float32((frame_idx - 1) / c.num_frames))
            with torch.no_grad():  # network_dynamic.py:261-265
                hash_2 = self._hash_dynamic(mod, x2, t2)
            plane_2 = self._planes(mod, torch.cat([x2, torch.full_like(tcol, t2)], -1), "dynamic")
        plane_d = 0.5 * plane_d + 0.25 * (plane_1 + plane_2)
        hash_d = 0.5 * hash_d + 0.25 * (hash_1 + hash_2)
        return torch.cat([plane_s, plane_d, hash_s, hash_d], dim=-1), flow

    def density(self, x, t, lidar):
        feats, _ = self.features(x, t, lidar)
        hdim = self.cfg.hidden
        h = mlp(self.p["sigma_net"], [(hdim, 128), (16, hdim)], feats, 1 + self.cfg.geo_feat_dim)
        return {"sigma": _TruncExp.apply(h[:, 0]), "geo_feat": h[:, 1:]}

    def color(self, d, geo_feat, lidar, mask=None):
        out_dim = 2 if lidar else 3
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], out_dim, dtype=torch.float32)
            if not mask.any():
                return rgbs
            d, geo_feat = d[mask], geo_feat[mask]
        hdim = self.cfg.hidden
        dn = (d + 1) / 2
        if lidar:
            enc = T._Frequency(3, {}).forward(dn)
            logits = torch.cat([enc, geo_feat], dim=-1)
            shapes = [(hdim, 96), (hdim, hdim), (16, hdim)]
            intensity = mlp(self.p["intensity_net"], shapes, logits, 1)
            raydrop = mlp(self.p["raydrop_net"], shapes, logits, 1)
            h = torch.cat([raydrop, intensity], dim=-1)
        else:
            enc = T.sh4(dn * 2.0 - 1.0)
            logits = torch.cat([enc, geo_feat], dim=-1)
            h = mlp(self.p["color_net"], [(hdim, 32), (hdim, hdim), (16, hdim)], logits, 3)
        h = torch.sigmoid(h)
        if mask is not None:
            rgbs[mask] = h
            return rgbs
        return h

    # ---------------------------------------------------------------- renderer
    def run(self, rays_o, rays_d, t, lidar, num_steps, nears=None, fars=None, noise=None, bg_color=1.0):
        """NeRFRenderer.run (renderer_dynamic.py:109-265).  noise: [N,S] in [0,1) or None
        (perturb=False).  Camera nears/fars come from near_far_from_aabb and are passed in."""
        c = self.cfg
        o, d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
        N = o.shape[0]
        if lidar:
            nears = torch.full((N,), c.min_near_lidar, dtype=torch.float32)
            fars = torch.full((N,), c.lidar_max_depth, dtype=torch.float32)
        nears, fars = nears.view(N, 1), fars.view(N, 1)
        z = torch.linspace(0.0, 1.0, num_steps).unsqueeze(0)
        z = nears + (fars - nears) * z
        sample_dist = (fars - nears) / num_steps
        if noise is not None:
            z = z + (noise - 0.5) * sample_dist
        xyz = o.unsqueeze(-2) + d.unsqueeze(-2) * z.unsqueeze(-1)
        xyz = torch.clamp(xyz, -c.bound, c.bound)
        dens = self.density(xyz.reshape(-1, 3), t, lidar)
        sigma = dens["sigma"].view(N, num_steps)
        deltas = torch.cat([z[:, 1:] - z[:, :-1], sample_dist * torch.ones_like(z[:, :1])], dim=-1)
        k = 2.0 if c.active_sensor else 1.0
        alphas = 1 - torch.exp(-k * deltas * c.density_scale * sigma)
        shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-15], dim=-1)
        weights = alphas * torch.cumprod(shifted, dim=-1)[:, :-1]
        mask = weights > 1e-4
        dirs = d.view(-1, 1, 3).expand_as(xyz)
        rgbs = self.color(dirs.reshape(-1, 3), dens["geo_feat"], lidar, mask=mask.reshape(-1))
        rgbs = rgbs.view(N, num_steps, -1)
        weights_sum = weights.sum(-1)
        depth = (weights * z).sum(-1)
        image = (weights.unsqueeze(-1) * rgbs).sum(-2)
        if not lidar:
            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        return dict(depth=depth, image=image, weights_sum=weights_sum, weights=weights, z_vals=z,
                    sigma=sigma)

    def render(self, rays_o, rays_d, t, lidar, num_steps, max_ray_batch=4096, nears=None, fars=None):
        """NeRFRenderer.render(staged=True) (renderer_dynamic.py:286-316): 4096-ray chunks."""
        N = rays_o.reshape(-1, 3).shape[0]
        depth = torch.empty(N)
        image = torch.empty(N, 2 if lidar else 3)
        for head in range(0, N, max_ray_batch):
            sl = slice(head, min(head + max_ray_batch, N))
            r = self.run(rays_o.reshape(-1, 3)[sl], rays_d.reshape(-1, 3)[sl], t, lidar, num_steps,
                         None if nears is None else nears[sl], None if fars is None else fars[sl])
            depth[sl], image[sl] = r["depth"], r["image"]
        return dict(depth=depth, image=image)
