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
