"""ctypes binding of libb200lap.so (C ABI declared in include/b200lap.h).

The library is the in-tree ``libb200lap.so`` next to this file, built by ``build.py`` with nvcc for
sm_100a.  There is no fallback: if the library is missing ``load()`` raises, and every compute
entry point fails loudly when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200lap.so")
if os.environ.get("B200LAP_LIB_VARIANT"):
    LIB_PATH = os.path.join(_HERE, "libb200lap_" + os.environ["B200LAP_LIB_VARIANT"] + ".so")    # experiment builds (tools/)
if os.environ.get("B200LAP_PROFILE_LIB") == "1":
    # measurement variant of the same sources (build.py --profile): SM-cycle counters in the solver trace; tools/ only
    LIB_PATH = os.path.join(_HERE, "libb200lap_prof.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_ll_p = ctypes.POINTER(ctypes.c_longlong)
vp = ctypes.c_void_p

ERR_CUDA, ERR_UNSUPPORTED, ERR_ARG = -100, -101, -102
ROW_FEAT_DIM = 21
TRACE_WORDS = 48
TRACE_NAMES = ("proj_triggers", "tight_edges", "greedy_matched", "took_fallback", "micro_bumps", "free_after_cr",
               "arr_iters", "aug_paths", "collect_calls", "relax_cols", "rc", "cyc_relax", "cyc_collect", "cyc_arr_scan",
               "cyc_arr_serial", "cyc_total", "collect_records", "cyc_collect_replay", "cyc_relax_replay", "relax_hits")

# name -> (restype, argtypes); the complete export list of include/b200lap.h
SIGNATURES = {
    "lapjv_seeded": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.c_double]),
    "b200lap_lapjv": (ctypes.c_int, [vp, ctypes.c_int, vp, vp]),
    "b200lap_lapjv_seeded_batch": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.c_double, vp, vp]),
    "b200lap_ctx_create": (ctypes.c_int, [ctypes.c_int, vp, ctypes.POINTER(vp)]),
    "b200lap_ctx_destroy": (None, [vp]),
    "b200lap_ctx_stream": (vp, [vp]),
    "b200lap_ctx_sync": (ctypes.c_int, [vp]),
    "b200lap_ctx_set_option": (ctypes.c_int, [vp, ctypes.c_char_p, ctypes.c_longlong]),
    "b200lap_ctx_launch_count": (ctypes.c_longlong, [vp]),
    "b200lap_ctx_feature_redo_rows": (ctypes.c_longlong, [vp]),
    "b200lap_last_error": (ctypes.c_char_p, []),
    "b200lap_device_count": (ctypes.c_int, []),
    "b200lap_dev_narrow": (ctypes.c_int, [vp, vp, ctypes.c_longlong, vp, vp]),
    "b200lap_dev_col_argmin": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]),
    "b200lap_dev_row_features": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp]),
    "b200lap_dev_onegnn_forward": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]),
    "b200lap_dev_min_trick": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]),
    "b200lap_dev_predict_duals": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]),
    "b200lap_dev_solve_seeded": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_double, vp, vp, vp, vp, vp]),
    "b200lap_dev_solve_cold": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]),
    "b200lap_dev_front_end": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_double, vp, vp, vp]),
    "b200lap_dev_pipeline": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, vp, vp, vp, vp, vp, vp]),
    "b200lap_model_create": (ctypes.c_int, [vp, vp, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]),
    "b200lap_model_destroy": (None, [vp]),
    "b200lap_compute_row_features": (ctypes.c_int, [vp, ctypes.c_int, vp]),
    "b200lap_pipeline_batch": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, vp, vp, vp, vp, vp, vp]),
    "b200lap_dev_project_feasible": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_double, vp]),
    "b200lap_dev_reduced_costs": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]),
    "b200lap_dev_bf_duals": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.POINTER(ctypes.c_int)]),
    "b200lap_project_feasible": (ctypes.c_int, [vp, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_double, vp]),
    "b200lap_reduce_costs": (ctypes.c_int, [vp, ctypes.c_int, vp, vp, ctypes.c_int, vp, vp]),
    "b200lap_ctx_lane_stream": (vp, [vp, ctypes.c_int]),
    "b200lap_ctx_last_lane": (ctypes.c_int, [vp]),
    "b200lap_ctx_join": (ctypes.c_int, [vp]),
    "b200lap_pipeline_batch_submit": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, vp, vp, vp, ctypes.POINTER(vp)]),
    "b200lap_pipeline_batch_wait": (ctypes.c_int, [vp]),
    "b200lap_host_narrow_config": (ctypes.c_longlong, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "b200lap_default_ctx": (vp, []),
}


class B200LapError(RuntimeError):
    pass


def bind(cdll):
    """Attach the declared signatures to an already-loaded library object."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)   # AttributeError here == the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    return cdll


_lib = None


def load():
    """Load the in-tree CUDA library (never anything else)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200LapError(
                f"{LIB_PATH} is missing: build it with `python build.py` (nvcc, sm_100a). "
                "libb200lap has no CPU fallback.")
        _lib = bind(ctypes.CDLL(LIB_PATH))
    return _lib


def last_error(lib=None) -> str:
    lib = lib or load()
    msg = lib.b200lap_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(code: int, what: str, lib=None) -> None:
    if code == 0:
        return
    raise B200LapError(f"{what} failed with code {code}: {last_error(lib)}")


def ptr(a):
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data
