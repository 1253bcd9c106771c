"""Drop-in for the dual-potential utilities of the reference's ``solvers/advanced_dual.py`` (lines 14-63), computed
by the B200 library: same names, arguments, return values and error behaviour, NumPy arrays in and out.

    project_feasible(C, u, v, max_rounds=50, tol=1e-12) -> (u, v)      /root/reference/solvers/advanced_dual.py:14-36
    reduce_costs(C, u, v, shift_nonneg=True) -> C'                     :39-53
    check_dual_feasible(C, u, v, tol=1e-8) -> True / AssertionError    :56-63

The sweeps are the solver's own kernels (row tightening and the feasibility predicate of ``k_front_end``, the column
sweep of the min-trick with binary64 row potentials, ``k_reduced_costs``); every value is a minimum or an element-wise
expression evaluated in binary64 exactly as the NumPy statements evaluate it, so results are bit-identical to the
reference's.  There is no CPU path: without a CUDA device the calls raise ``B200LapError``.
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import numpy as np

from b200lap import _lib


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=float), dtype=np.float64)


def _square(C: np.ndarray) -> int:
    if C.ndim != 2 or C.shape[0] != C.shape[1]:
        raise ValueError("C must be a square matrix")
    return int(C.shape[0])


def project_feasible(C: np.ndarray, u: np.ndarray, v: np.ndarray, max_rounds: int = 50, tol: float = 1e-12) -> Tuple[np.ndarray, np.ndarray]:
    """Iteratively tighten (u, v) until min(C - u - v) >= -tol or the rounds are exhausted."""
    lib = _lib.load()
    C = _f64(C)
    n = _square(C)
    u = _f64(u).copy()
    v = _f64(v).copy()
    if u.shape != (n,) or v.shape != (n,):
        raise ValueError("u/v sizes must match C")
    rounds = ctypes.c_int(0)
    _lib.check(lib.b200lap_project_feasible(C.ctypes.data, n, u.ctypes.data, v.ctypes.data, max(1, int(max_rounds)), float(tol),
                                            ctypes.addressof(rounds)), "b200lap_project_feasible", lib)
    return u, v


def reduce_costs(C: np.ndarray, u: np.ndarray, v: np.ndarray, shift_nonneg: bool = True) -> np.ndarray:
    """C' = C - u 1^T - 1 v^T; with ``shift_nonneg`` the minimum is subtracted when it is negative."""
    lib = _lib.load()
    C = _f64(C)
    n = _square(C)
    u, v = _f64(u), _f64(v)
    out = np.empty((n, n), dtype=np.float64)
    mn = ctypes.c_double(0.0)
    _lib.check(lib.b200lap_reduce_costs(C.ctypes.data, n, u.ctypes.data, v.ctypes.data, int(bool(shift_nonneg)), out.ctypes.data,
                                        ctypes.addressof(mn)), "b200lap_reduce_costs", lib)
    return out


def min_reduced_cost(C: np.ndarray, u: np.ndarray, v: np.ndarray) -> float:
    """min_ij (C_ij - u_i - v_j) without materialising the matrix (what ``check_dual_feasible`` compares with -tol)."""
    lib = _lib.load()
    C = _f64(C)
    n = _square(C)
    u, v = _f64(u), _f64(v)
    mn = ctypes.c_double(0.0)
    _lib.check(lib.b200lap_reduce_costs(C.ctypes.data, n, u.ctypes.data, v.ctypes.data, 0, None, ctypes.addressof(mn)),
               "b200lap_reduce_costs", lib)
    return float(mn.value)


def check_dual_feasible(C: np.ndarray, u: np.ndarray, v: np.ndarray, tol: float = 1e-8) -> bool:
    """Assert dual feasibility only: r_ij = C_ij - u_i - v_j >= -tol for all i, j."""
    mn = min_reduced_cost(C, u, v)
    if mn < -tol:
        raise AssertionError(f"Dual infeasible: min reduced cost {mn:.3e} < -tol")
    return True
