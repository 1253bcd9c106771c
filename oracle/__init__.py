"""oracle/ -- TEST INFRASTRUCTURE ONLY (the checker, never the product).

CPU restatements of the reference hot path used to prove parity of the CUDA
implementation.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package; nothing under ``gnn-accelerated-lap-warm-start-pipeline_b200/`` does.

Contents
  jv_port.c       our C restatement of lapjv_seeded / lapjv_internal
                  (-> libjvport.so; parity PINNED against oracle/_ref and the
                  reference's known-answer vectors, see its header)
  _ref/           the UNMODIFIED reference solver compiled from
                  /root/reference/LAP/_lapjv_cpp/{lapjv_seeded,lapjv}.cpp
                  (git-ignored build output; travels to the GPU box)
  features_np.py  numpy restatement of gnn/features.py:compute_row_features
  onegnn_np.py    numpy restatement of gnn/one_gnn.py:OneGNN.forward
  pipeline_np.py  the CPU dtype path of scripts/gnn_benchmark.py:226-262,289

This module holds the ctypes bindings for the two shared libraries.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "libjvport.so")
_REF_SO = os.path.join(_HERE, "_ref", "libreflap.so")

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_ll_p = ctypes.POINTER(ctypes.c_longlong)


class Trace(ctypes.Structure):
    """Mirror of ``struct jvp_trace`` in jv_port.c."""

    _fields_ = [
        ("proj_triggers", ctypes.c_int64),
        ("tight_edges", ctypes.c_int64),
        ("greedy_matched", ctypes.c_int64),
        ("took_fallback", ctypes.c_int64),
        ("micro_bumps", ctypes.c_int64),
        ("free_after_cr", ctypes.c_int64),
        ("arr_iters", ctypes.c_int64),
        ("aug_paths", ctypes.c_int64),
        ("collect_calls", ctypes.c_int64),
        ("relax_cols", ctypes.c_int64),
        ("rc", ctypes.c_int64),
        ("path_log", ctypes.POINTER(ctypes.c_int32)),
        ("path_log_cap", ctypes.c_int64),
    ]

    def as_dict(self) -> dict:
        return {k: int(getattr(self, k)) for k, _ in self._fields_ if k not in ("path_log", "path_log_cap")}


def build(force: bool = False) -> None:
    """Compile the checkers (``make -C oracle``).  Building is not using."""
    need_port = force or not os.path.exists(_PORT_SO) or (
        os.path.getmtime(_PORT_SO) < os.path.getmtime(os.path.join(_HERE, "jv_port.c")))
    if need_port:
        subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/LAP/_lapjv_cpp") and (force or not os.path.exists(_REF_SO)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        build()
        lib = ctypes.CDLL(_PORT_SO)
        lib.jvp_lapjv_seeded.restype = ctypes.c_int
        lib.jvp_lapjv_seeded.argtypes = [_c_double_p, ctypes.c_int, ctypes.c_int, _c_ll_p, _c_ll_p,
                                         _c_double_p, _c_double_p, ctypes.c_double, ctypes.POINTER(Trace)]
        lib.jvp_lapjv.restype = ctypes.c_int
        lib.jvp_lapjv.argtypes = [_c_double_p, ctypes.c_int, _c_int_p, _c_int_p, ctypes.POINTER(Trace)]
        lib.jvp_front_end.restype = ctypes.c_int
        lib.jvp_front_end.argtypes = [_c_double_p, ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_double,
                                      _c_double_p, _c_double_p, _c_int_p, _c_int_p, ctypes.POINTER(Trace)]
        lib.jvp_lapjv_duals.restype = ctypes.c_int
        lib.jvp_lapjv_duals.argtypes = [_c_double_p, ctypes.c_int, _c_int_p, _c_int_p, _c_double_p]
        _port = lib
    return _port


def ref_available() -> bool:
    if not os.path.exists(_REF_SO):
        try:
            build()
        except Exception:
            return False
    return os.path.exists(_REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref/libreflap.so is missing (run `make -C oracle ref` where /root/reference exists)")
        lib = ctypes.CDLL(_REF_SO)
        lib.lapjv_seeded.restype = ctypes.c_int
        lib.lapjv_seeded.argtypes = [_c_double_p, ctypes.c_int, ctypes.c_int, _c_ll_p, _c_ll_p,
                                     _c_double_p, _c_double_p, ctypes.c_double]
        lib.ref_lapjv_internal.restype = ctypes.c_int
        lib.ref_lapjv_internal.argtypes = [_c_double_p, ctypes.c_int, _c_int_p, _c_int_p]
        _ref = lib
    return _ref


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_c_double_p)


def _seeded_call(fn, C, u, v, eps, trace: Optional[Trace]):
    """Restates the Python binding LAP/lap/_seeded_jv.pyx:14-31 around a C entry point."""
    C = _f64(C)
    u = _f64(u)
    v = _f64(v)
    if C.ndim != 2 or u.ndim != 1 or v.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions")
    n, m = C.shape
    if u.shape[0] != n or v.shape[0] != m:
        raise ValueError("u/v sizes must match C")
    x = np.full((n,), -1, dtype=np.int64)
    y = np.full((m,), -1, dtype=np.int64)
    args = [_dp(C), n, m, x.ctypes.data_as(_c_ll_p), y.ctypes.data_as(_c_ll_p), _dp(u), _dp(v), float(eps)]
    if trace is not None:
        args.append(ctypes.byref(trace))
    rc = fn(*args)
    if rc != 0:
        if rc == -3:
            raise ValueError("Infeasible seed potentials: C - u - v has negatives")
        raise RuntimeError(f"lapjv_seeded internal error (code {rc})")
    cost = float(np.sum(C[np.arange(n), x]))
    return x, y, cost


def ref_lapjv_seeded(C, u, v, eps: float = 1e-12):
    """The unmodified reference ``lapjv_seeded`` (oracle/_ref)."""
    return _seeded_call(ref_lib().lapjv_seeded, C, u, v, eps, None)


def port_lapjv_seeded(C, u, v, eps: float = 1e-12, trace: Optional[Trace] = None):
    """Our C restatement; ``trace`` (a :class:`Trace`) receives the phase counters."""
    if trace is None:
        trace = Trace()
    return _seeded_call(port_lib().jvp_lapjv_seeded, C, u, v, eps, trace)


def _cold_call(fn, C, with_trace: Optional[Trace]):
    C = _f64(C)
    n = C.shape[0]
    assert C.shape == (n, n)
    x = np.empty((n,), dtype=np.int32)
    y = np.empty((n,), dtype=np.int32)
    args = [_dp(C), n, x.ctypes.data_as(_c_int_p), y.ctypes.data_as(_c_int_p)]
    if with_trace is not None:
        args.append(ctypes.byref(with_trace))
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"lapjv_internal returned {rc}")
    return x, y


def ref_lapjv_internal(C) -> Tuple[np.ndarray, np.ndarray]:
    """The unmodified reference ``lapjv_internal`` on a square matrix -> (x, y) int32."""
    return _cold_call(ref_lib().ref_lapjv_internal, C, None)


def port_lapjv_internal(C, trace: Optional[Trace] = None) -> Tuple[np.ndarray, np.ndarray]:
    if trace is None:
        trace = Trace()
    return _cold_call(port_lib().jvp_lapjv, C, trace)


def port_optimal_duals(C):
    """Cold solve -> (x, y, u, v) with (u, v) the optimal duals JV ends with
    (v final, u_i = C[i, x_i] - v[x_i])."""
    C = _f64(C)
    n = C.shape[0]
    x = np.empty(n, dtype=np.int32)
    y = np.empty(n, dtype=np.int32)
    v = np.empty(n)
    rc = port_lib().jvp_lapjv_duals(_dp(C), n, x.ctypes.data_as(_c_int_p), y.ctypes.data_as(_c_int_p), _dp(v))
    if rc != 0:
        raise RuntimeError(f"jvp_lapjv_duals returned {rc}")
    u = C[np.arange(n), x] - v[x]
    return x, y, u, v


def port_front_end(C, u_seed, v_seed, eps: float = 1e-12):
    """Projection + row tightening + greedy + tight count -> (rc, u, v, x, y, trace-dict)."""
    C = _f64(C)
    n = C.shape[0]
    us, vs = _f64(u_seed), _f64(v_seed)
    u = np.empty(n)
    v = np.empty(n)
    x = np.empty(n, dtype=np.int32)
    y = np.empty(n, dtype=np.int32)
    t = Trace()
    rc = port_lib().jvp_front_end(_dp(C), n, _dp(us), _dp(vs), float(eps), _dp(u), _dp(v),
                                  x.ctypes.data_as(_c_int_p), y.ctypes.data_as(_c_int_p), ctypes.byref(t))
    return rc, u, v, x, y, t.as_dict()


def lapjv_py(cost, extend_cost: bool = False, cost_limit: float = np.inf, return_cost: bool = True,
             *, use_ref: bool = False):
    """Restates the Python binding LAP/_lapjv_cpp/_lapjv.pyx:38-129 around a cold solve."""
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
    x_c, y_c = (ref_lapjv_internal if use_ref else port_lapjv_internal)(cost_c)
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
