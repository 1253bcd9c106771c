"""Thin helpers that drive the C ABI with plain numpy buffers (used by the emulator-backed logic tests,
where "device" pointers are host pointers, and by the GPU tests through pinned-free host entry points)."""
import ctypes

import numpy as np

from b200lap import _lib
from b200lap.runtime import pack_state_dict

TRACE = _lib.TRACE_NAMES


def seeded(lib, C, u, v, eps=1e-12):
    """lapjv_seeded_batch with batch 1 -> (rc, x, y, trace dict)."""
    C = np.ascontiguousarray(C, dtype=np.float64)
    n = C.shape[0]
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    x = np.full(n, -1, np.int64)
    y = np.full(n, -1, np.int64)
    rc = np.zeros(1, np.int32)
    tr = np.zeros(_lib.TRACE_WORDS, np.int64)
    r = lib.b200lap_lapjv_seeded_batch(C.ctypes.data, 1, n, x.ctypes.data, y.ctypes.data, u.ctypes.data, v.ctypes.data,
                                       float(eps), rc.ctypes.data, tr.ctypes.data)
    return (r or int(rc[0])), x, y, {k: int(tr[i]) for i, k in enumerate(TRACE)}


def seeded_dropin(lib, C, u, v, eps=1e-12):
    C = np.ascontiguousarray(C, dtype=np.float64)
    n, m = C.shape
    x = np.full(n, -1, np.int64)
    y = np.full(m, -1, np.int64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    rc = lib.lapjv_seeded(C.ctypes.data, n, m, x.ctypes.data, y.ctypes.data, u.ctypes.data, v.ctypes.data, float(eps))
    return rc, x, y


def cold(lib, C):
    C = np.ascontiguousarray(C, dtype=np.float64)
    n = C.shape[0]
    x = np.zeros(n, np.int32)
    y = np.zeros(n, np.int32)
    rc = lib.b200lap_lapjv(C.ctypes.data, n, x.ctypes.data, y.ctypes.data)
    return rc, x, y


def make_model(lib, ctx, sd, topk):
    blob, in_dim, hidden, layers = pack_state_dict(sd)
    h = ctypes.c_void_p()
    rc = lib.b200lap_model_create(ctx, blob.ctypes.data, blob.size, in_dim, hidden, layers, topk, ctypes.byref(h))
    assert rc == 0, _lib.last_error(lib)
    return h


def trace_matches(oracle_trace: dict, got: dict):
    keys = ("proj_triggers", "tight_edges", "greedy_matched", "took_fallback", "micro_bumps", "free_after_cr",
            "arr_iters", "aug_paths", "collect_calls", "relax_cols")
    return [(k, oracle_trace[k], got[k]) for k in keys if oracle_trace[k] != got[k]]
