"""``LAPSolver`` / ``SeededLAPSolver`` -- same classes, methods and return conventions as
/root/reference/solvers/lap_solver.py:33-105, bound to the B200 ``lap`` drop-in.

As in the reference, ``SeededLAPSolver.solve`` returns ``(x, y, cost)`` under the names
``(rows, cols, cost)``: the row->column map and the column->row map, not index pairs.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

import lap


def _resolve_seeded_api():
    seeded = getattr(lap, "lapjv_seeded", None)
    if seeded is not None:
        return seeded
    try:
        from lap._seeded_jv import lapjv_seeded  # type: ignore
        return lapjv_seeded
    except Exception:
        return None


_LAPJV_SEEDED = _resolve_seeded_api()


class LAPSolver:
    """Unseeded JV (``lap.lapjv``)."""

    def __init__(self):
        self.name = "LAP"

    def solve(self, C: np.ndarray) -> Tuple[np.ndarray, np.ndarray, float]:
        C = np.asarray(C, dtype=np.float64)
        n = C.shape[0]
        _, x, _ = lap.lapjv(C, extend_cost=False)
        rows = np.arange(n, dtype=np.int64)
        cols = np.asarray(x, dtype=np.int64)
        cost = sum(C[i, cols[i]] for i in range(n) if cols[i] >= 0)
        return rows, cols, float(cost)

    def __call__(self, C):
        return self.solve(C)


class SeededLAPSolver:
    """Seeded JV (``lap.lapjv_seeded``) warm-started with dual potentials."""

    def __init__(self):
        self.name = "SeededLAP"
        if _LAPJV_SEEDED is None:
            raise ImportError("lap.lapjv_seeded is not available")

    def solve(self, C: np.ndarray, u: np.ndarray, v: np.ndarray) -> Tuple[np.ndarray, np.ndarray, float]:
        C = np.ascontiguousarray(np.asarray(C, dtype=np.float64))
        u = np.ascontiguousarray(np.asarray(u, dtype=np.float64))
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
        rows, cols, cost = _LAPJV_SEEDED(C, u, v)
        return np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64), float(cost)

    def __call__(self, C, u, v):
        return self.solve(C, u, v)
