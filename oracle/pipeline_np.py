"""oracle/pipeline_np.py -- TEST INFRASTRUCTURE ONLY (the checker, never the product).

The dtype path of the reference's CPU inference glue, restated:
``/root/reference/scripts/gnn_benchmark.py:226-262,289`` (GNNPredictor.predict, CPU branch)
followed by ``/root/reference/solvers/lap_solver.py:81-101`` (SeededLAPSolver.solve).

  features  : binary64 statistics of C -> f32[n,21]
  OneGNN    : binary32, ``cost`` = C cast to binary32, all-true mask
  min-trick : v_j = min_i (C64_ij - float64(u32_i)), binary64
  solve     : lapjv_seeded(C64, float64(u), v)

Parity status: PINNED through its parts (features_np, onegnn_np, jv_port.c) and, end to
end, against tests/golden/dense_golden.npz.
"""
from __future__ import annotations

import numpy as np

from . import features_np, onegnn_np


def min_trick(C, u) -> np.ndarray:
    """gnn_benchmark.py:262: column potentials from row potentials, binary64."""
    C = np.asarray(C, dtype=np.float64)
    return np.min(C - np.asarray(u, dtype=np.float64)[:, None], axis=0)


def predict_duals(C, sd, topk: int = 16):
    """gnn_benchmark.py:226-262,289 -> (u f64[n] holding binary32 values, v f64[n])."""
    C = np.asarray(C, dtype=np.float64)
    feat = features_np.row_features(C)
    mask = np.ones(C.shape[0], dtype=bool)
    u32 = onegnn_np.forward(sd, feat, cost=C.astype(np.float32), mask=mask, topk=topk)
    v = min_trick(C, u32)
    return u32.astype(np.float64), v.astype(np.float64)


def solve(C, sd, topk: int = 16, use_ref: bool = False):
    """features -> OneGNN -> min-trick -> seeded JV; returns (x, y, cost, u, v)."""
    import oracle
    u, v = predict_duals(C, sd, topk)
    fn = oracle.ref_lapjv_seeded if use_ref else oracle.port_lapjv_seeded
    x, y, cost = fn(np.asarray(C, dtype=np.float64), u, v)
    return x, y, cost, u, v


# -- numpy statements of the front-end sweeps (solvers/advanced_dual.py) ---------------------

def project_feasible(C, u, v, max_rounds: int = 50, tol: float = 1e-12):
    """solvers/advanced_dual.py:14-36 (Jacobi clamp rounds until min reduced cost >= -tol)."""
    C = np.asarray(C, dtype=float)
    u = np.array(u, dtype=float)
    v = np.array(v, dtype=float)
    for _ in range(max(1, int(max_rounds))):
        u = np.minimum(u, (C - v[None, :]).min(axis=1))
        v = np.minimum(v, (C - u[:, None]).min(axis=0))
        if (C - u[:, None] - v[None, :]).min() >= -tol:
            break
    return u, v


def reduce_costs(C, u, v, shift_nonneg: bool = True):
    """solvers/advanced_dual.py:39-53."""
    R = np.asarray(C, dtype=float) - np.asarray(u)[:, None] - np.asarray(v)[None, :]
    if shift_nonneg:
        m = R.min()
        if m < 0:
            R = R - m
    return np.ascontiguousarray(R, dtype=np.float64)


def tight_edge_count(C, u, v, tol: float = 1e-9) -> int:
    """lapjv_seeded.cpp:105-113 in numpy: #{(i,j): |(C_ij - u_i) - v_j| <= tol}."""
    R = (np.asarray(C, dtype=float) - np.asarray(u)[:, None]) - np.asarray(v)[None, :]
    return int((np.abs(R) <= tol).sum())
