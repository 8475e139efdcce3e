"""Shared helpers for field parity tests (oracle side)."""
import importlib

import numpy as np
import torch

S = importlib.import_module("selfsupervised-nvsf_b200.synth")
_cache = {}


def oracle_config(density_scale=1.0, **kw):
    from oracle.field_oracle import FieldConfig
    return FieldConfig(bound=S.BOUND, num_frames=S.NUM_FRAMES, time_resolution=S.TIME_RESOLUTION,
                       min_near=S.MIN_NEAR, min_near_lidar=S.MIN_NEAR_LIDAR, lidar_max_depth=S.LIDAR_MAX_DEPTH,
                       density_scale=density_scale, **kw)


def oracle_params(seed=0, style="trained"):
    """Full-size parameters (62 M floats) from the numpy-seeded initialiser; cached per process."""
    key = (seed, style)
    if key not in _cache:
        from oracle import field_init
        _cache[key] = field_init.make_params(oracle_config(), seed=seed, style=style)
    return _cache[key]


def rel_err(a, b, floor=1e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float((np.abs(a - b) / (np.abs(b) + floor)).max())
