"""``lap`` -- drop-in for the reference's forked LAP package on the seeded hot path.

Same call signatures, argument checks, exceptions and return values as
/root/reference/LAP/lap/_seeded_jv.pyx:14-31 (``lapjv_seeded``) and
/root/reference/LAP/_lapjv_cpp/_lapjv.pyx:38-129 (``lapjv``); the work happens in the sm_100a kernels
behind libb200lap.so (C symbols ``lapjv_seeded`` / ``b200lap_lapjv``).  There is no CPU path: without
the CUDA library or a device both functions raise.  ``lapmod`` (the sparse solver) is out of scope.
"""
from __future__ import annotations

import numpy as np

__version__ = "0.5.12+b200"

LARGE = 1000000.0    # /root/reference/LAP/_lapjv_cpp/lapjv.h:4
FP_1, FP_2, FP_DYNAMIC = 1, 2, 3

__all__ = ["lapjv", "lapjv_seeded", "LARGE", "FP_1", "FP_2", "FP_DYNAMIC"]


def _lib():
    from b200lap import _lib as L
    return L.load()


def _buffer(a, ndim: int, name: str) -> np.ndarray:
    # the Cython signature is ndarray[double, ndim, mode='c']: wrong dtype/ndim/layout -> ValueError
    if not isinstance(a, np.ndarray):
        raise TypeError(f"Argument '{name}' has incorrect type (expected numpy.ndarray, got {type(a).__name__})")
    if a.dtype != np.float64:
        raise ValueError(f"Buffer dtype mismatch, expected 'double' but got '{a.dtype}'")
    if a.ndim != ndim:
        raise ValueError(f"Buffer has wrong number of dimensions (expected {ndim}, got {a.ndim})")
    if not a.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    return a


def lapjv_seeded(C, u, v, eps: float = 1e-12):
    """Seeded Jonker-Volgenant: (x int64[n], y int64[m], cost float) from (C, u, v) float64 arrays."""
    C = _buffer(C, 2, "C")
    u = _buffer(u, 1, "u")
    v = _buffer(v, 1, "v")
    n, m = C.shape
    if u.shape[0] != n or v.shape[0] != m:
        raise ValueError("u/v sizes must match C")
    x = np.full((n,), -1, dtype=np.int64)
    y = np.full((m,), -1, dtype=np.int64)
    ret = _lib().lapjv_seeded(C.ctypes.data, n, m, x.ctypes.data, y.ctypes.data, u.ctypes.data, v.ctypes.data, float(eps))
    if ret != 0:
        if ret == -3:
            raise ValueError("Infeasible seed potentials: C - u - v has negatives")
        raise RuntimeError(f"lapjv_seeded internal error (code {ret})")
    cost = float(np.sum(C[np.arange(n), x]))
    return x, y, cost


def lapjv(cost, extend_cost: bool = False, cost_limit: float = np.inf, return_cost: bool = True):
    """Cold Jonker-Volgenant with the reference's padding rules: (opt, x, y) or (x, y)."""
    cost = np.asarray(cost)
    if cost.ndim != 2:
        raise ValueError("2-dimensional array expected")
    cost_c = np.ascontiguousarray(cost, dtype=np.double)
    n_rows, n_cols = cost_c.shape
    n = 0
    if n_rows == n_cols:
        n = n_rows
    elif not extend_cost:
        raise ValueError("Square cost array expected. If cost is intentionally non-square, pass extend_cost=True.")
    if cost_limit < np.inf:
        n = n_rows + n_cols
        ext = np.empty((n, n), dtype=np.double)
        ext[:] = cost_limit / 2.0
        ext[n_rows:, n_cols:] = 0
        ext[:n_rows, :n_cols] = cost_c
        cost_c = ext
    elif extend_cost:
        n = max(n_rows, n_cols)
        ext = np.zeros((n, n), dtype=np.double)
        ext[:n_rows, :n_cols] = cost_c
        cost_c = ext
    x_c = np.empty((n,), dtype=np.int32)
    y_c = np.empty((n,), dtype=np.int32)
    ret = _lib().b200lap_lapjv(cost_c.ctypes.data, n, x_c.ctypes.data, y_c.ctypes.data)
    if ret != 0:
        if ret == -1:
            raise MemoryError("Out of memory.")
        raise RuntimeError("Unknown error (lapjv_internal returned %d)." % ret)
    opt = np.nan
    if cost_limit < np.inf or extend_cost:
        x_c[x_c >= n_cols] = -1
        y_c[y_c >= n_rows] = -1
        x_c = x_c[:n_rows]
        y_c = y_c[:n_cols]
        if return_cost:
            opt = cost_c[np.nonzero(x_c != -1)[0], x_c[x_c != -1]].sum()
    elif return_cost:
        opt = cost_c[np.arange(n_rows), x_c].sum()
    if return_cost:
        return opt, x_c, y_c
    return x_c, y_c
