"""TEST INFRASTRUCTURE.  numpy front-end of the Chamfer part of the C oracle
(oracle/raymarching_oracle.c, restating nvsf/nerf/chamfer3D/chamfer3D.cu); same call shape as the
reference module chamfer_3DDist (dist_chamfer_3D.py:86-95): (xyz1 [B,N,3], xyz2 [B,M,3]) ->
dist1 [B,N], dist2 [B,M], idx1 [B,N], idx2 [B,M]."""
import ctypes

import numpy as np

from .raymarching_oracle import lib

_P = ctypes.c_void_p
_U = ctypes.c_uint32


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def chamfer_forward(xyz1, xyz2):
    a, b = _f(xyz1), _f(xyz2)
    B, n, m = a.shape[0], a.shape[1], b.shape[1]
    L = lib()
    L.oracle_chamfer_nn.restype = None
    d1, d2 = np.zeros((B, n), np.float32), np.zeros((B, m), np.float32)
    i1, i2 = np.zeros((B, n), np.int32), np.zeros((B, m), np.int32)
    L.oracle_chamfer_nn(_p(a), _U(n), _p(b), _U(m), _U(B), _p(d1), _p(i1))
    L.oracle_chamfer_nn(_p(b), _U(m), _p(a), _U(n), _U(B), _p(d2), _p(i2))
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, grad_dist1, grad_dist2, idx1, idx2):
    a, b = _f(xyz1), _f(xyz2)
    B, n, m = a.shape[0], a.shape[1], b.shape[1]
    L = lib()
    L.oracle_chamfer_grad.restype = None
    g1, g2 = np.zeros_like(a), np.zeros_like(b)
    i1, i2 = np.ascontiguousarray(idx1, np.int32), np.ascontiguousarray(idx2, np.int32)
    L.oracle_chamfer_grad(_p(a), _U(n), _p(b), _U(m), _U(B), _p(_f(grad_dist1)), _p(i1), _p(g1), _p(g2))
    L.oracle_chamfer_grad(_p(b), _U(m), _p(a), _U(n), _U(B), _p(_f(grad_dist2)), _p(i2), _p(g2), _p(g1))
    return g1, g2
