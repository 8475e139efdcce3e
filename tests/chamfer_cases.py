"""Seeded Chamfer inputs shared by the golden generator (oracle/make_golden_chamfer.py), the CPU
oracle test and the GPU parity test."""
import numpy as np

# name -> (batch, n, m, seed); sizes straddle the reference kernel's 512-point shared-memory blocks
# (chamfer3D.cu:16) and its 4-way unrolled / tail loops (:32-127)
GOLDEN_CASES = {
    "small": (2, 1000, 777, 1),
    "blocks": (1, 1024, 512, 2),
    "tail": (1, 513, 515, 3),
    "tiny": (1, 5, 3, 4),
    "lidar": (1, 4096, 4096, 5),     # the trainer's call: predicted vs gt points of one ray batch
    "ties": (1, 300, 256, 6),        # duplicated targets: the first index must win
}


def case(name):
    b, n, m, seed = GOLDEN_CASES[name]
    rng = np.random.default_rng(seed)
    a = (rng.random((b, n, 3), dtype=np.float32) * 2 - 1).astype(np.float32)
    t = (rng.random((b, m, 3), dtype=np.float32) * 2 - 1).astype(np.float32)
    if name == "ties":
        t[:, m // 2:] = t[:, :m - m // 2]          # every target appears twice
        a[:, :50] = t[:, 100:150]                  # exact hits (distance 0)
    if name == "lidar":                            # rays_d * depth / scale (trainer.py:230-231)
        d = a / np.linalg.norm(a, axis=-1, keepdims=True)
        a = (d * rng.random((b, n, 1), dtype=np.float32) * 80).astype(np.float32)
        t = (d * rng.random((b, n, 1), dtype=np.float32) * 80).astype(np.float32)
    g1 = rng.random((b, n), dtype=np.float32)
    g2 = rng.random((b, m), dtype=np.float32)
    return a, t, g1, g2
