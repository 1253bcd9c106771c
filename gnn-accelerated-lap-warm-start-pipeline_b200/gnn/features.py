"""``compute_row_features`` -- same signature and output as /root/reference/gnn/features.py:161-243.

The statistics are computed by the row-resident sm_100a sweep (csrc/features.cuh) through the C symbol
``b200lap_compute_row_features``; this module only checks arguments and moves the buffers.
"""
from __future__ import annotations

import numpy as np

ROW_FEAT_DIM = 21
TAU = 1e-3                # /root/reference/gnn/features.py:16-18
POS_FREQS = (1, 2, 4, 8)
EPS = 1e-9


def compute_row_features(C) -> np.ndarray:
    """21-D row features of a square cost matrix -> float32[n, 21]."""
    C = np.ascontiguousarray(np.asarray(C, dtype=np.float64))
    if C.ndim != 2:
        raise ValueError("C must be 2-dimensional")
    n = C.shape[0]
    if n == 0:
        return np.zeros((0, 0), dtype=np.float32)
    if C.shape[1] != n:
        raise ValueError("compute_row_features on the device path expects a square matrix")
    from b200lap import _lib as L
    lib = L.load()
    feat = np.empty((n, ROW_FEAT_DIM), dtype=np.float32)
    L.check(lib.b200lap_compute_row_features(C.ctypes.data, n, feat.ctypes.data), "b200lap_compute_row_features", lib)
    return feat


def compute_row_features_torch(cost):
    """Same signature and definitions as /root/reference/gnn/features.py:246-351: a CUDA tensor [n, n] in, the 21-D row
    features as a float32 CUDA tensor [n, 21] out, nothing leaves the device.  (Differences from ``compute_row_features``
    that the reference's torch variant has and this mirror keeps: unbiased ``row_std`` / ``k_std``, ``is_col_best`` counts
    only the first row attaining each column minimum, the near-best threshold is multiplied in binary32.)"""
    import torch
    import b200lap
    if not isinstance(cost, torch.Tensor):
        raise TypeError("compute_row_features_torch expects a torch.Tensor")
    C = cost.float()
    if C.dim() != 2:
        raise ValueError("cost must be 2-dimensional")
    n, m = C.shape
    if n == 0:
        return torch.zeros((0, 0), dtype=torch.float32, device=C.device)
    if n != m:
        raise ValueError("compute_row_features_torch on the device path expects a square matrix")
    if not C.is_cuda:
        raise RuntimeError("compute_row_features_torch runs on the GPU: pass a CUDA tensor (there is no CPU path)")
    ctx = b200lap.default_context(C.device.index or 0)
    feat = ctx.row_features_torch(C.contiguous())
    ctx.sync()
    return feat[0]
