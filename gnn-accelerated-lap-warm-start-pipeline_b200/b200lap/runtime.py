"""Host-side runtime over the C ABI: contexts, packed models and the device-pointer entry points.

PyTorch is used here for device memory and stream bookkeeping only (tensors are allocation handles whose
``data_ptr()`` goes straight into the C ABI); every computation is a hand-written kernel in
``libb200lap.so``.  Nothing in this module has a CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Mapping, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import B200LapError, TRACE_NAMES, TRACE_WORDS, ROW_FEAT_DIM, check, ptr

STATE_DICT_HEAD = ("input_proj.0.weight", "input_proj.0.bias", "input_proj.2.weight", "input_proj.2.bias")
STATE_DICT_TAIL = ("pre_out.weight", "pre_out.bias", "row_out.0.weight", "row_out.0.bias", "row_out.3.weight",
                   "row_out.3.bias", "edge_mlp.0.weight", "edge_mlp.0.bias", "edge_mlp.2.weight", "edge_mlp.2.bias",
                   "message_norm.weight", "message_norm.bias")


def state_dict_order(layers: int):
    """Parameter order of the packed blob == registration order of /root/reference/gnn/one_gnn.py:64-87."""
    keys = list(STATE_DICT_HEAD)
    for l in range(layers):
        keys += [f"blocks.{l}.fc1.weight", f"blocks.{l}.fc1.bias", f"blocks.{l}.fc2.weight", f"blocks.{l}.fc2.bias",
                 f"blocks.{l}.norm.weight", f"blocks.{l}.norm.bias"]
    return keys + list(STATE_DICT_TAIL)


def count_layers(sd: Mapping) -> int:
    k = 0
    while f"blocks.{k}.fc1.weight" in sd:
        k += 1
    return k


def pack_state_dict(sd: Mapping) -> Tuple[np.ndarray, int, int, int]:
    """state_dict (numpy arrays or torch tensors) -> (float32 blob, in_dim, hidden, layers)."""
    def arr(t):
        if hasattr(t, "detach"):
            t = t.detach().cpu().numpy()
        return np.ascontiguousarray(t, dtype=np.float32).ravel()
    layers = count_layers(sd)
    w0 = sd["input_proj.0.weight"]
    hidden, in_dim = int(w0.shape[0]), int(w0.shape[1])
    blob = np.concatenate([arr(sd[k]) for k in state_dict_order(layers)])
    return blob, in_dim, hidden, layers


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise B200LapError("no CUDA device is visible: b200lap has no CPU path")
    return torch


def _ordered(fn):
    """Order the context's stream after torch's current stream before the call and torch's current
    stream after the context's stream afterwards, so tensors produced by torch ops are complete when
    our kernels read them and our results are complete when torch ops read them."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        torch = _torch()
        cur = torch.cuda.current_stream(self.device)
        mine = self.torch_stream()
        if cur.cuda_stream != mine.cuda_stream:
            mine.wait_stream(cur)
        with torch.cuda.device(self.device):
            out = fn(self, *args, **kwargs)
        if cur.cuda_stream != mine.cuda_stream:
            cur.wait_stream(mine)
        return out
    return wrapper


class Context:
    """One (device, stream) pair plus cached device workspaces (b200lap_ctx)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        check(self.lib.b200lap_ctx_create(int(device), stream, ctypes.byref(h)), "b200lap_ctx_create", self.lib)
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.b200lap_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.b200lap_ctx_stream(self.handle) or 0)

    def torch_stream(self):
        torch = _torch()
        if getattr(self, "_tstream", None) is None:
            self._tstream = torch.cuda.ExternalStream(self.stream_ptr, device=self.device)
        return self._tstream

    def sync(self):
        check(self.lib.b200lap_ctx_sync(self.handle), "b200lap_ctx_sync", self.lib)

    def set_option(self, key: str, value: int):
        check(self.lib.b200lap_ctx_set_option(self.handle, key.encode(), int(value)), "b200lap_ctx_set_option", self.lib)

    @property
    def launches(self) -> int:
        return int(self.lib.b200lap_ctx_launch_count(self.handle))

    def feature_redo_rows(self):
        """Rows of the last row-feature call that the fast kernel handed to the exact fall-back (diagnostic)."""
        return int(self.lib.b200lap_ctx_feature_redo_rows(self.handle))

    # ---- device-pointer entry points (arguments are torch CUDA tensors on this context's device) ----
    def _empty(self, shape, dtype):
        torch = _torch()
        return torch.empty(shape, dtype=dtype, device=f"cuda:{self.device}")

    @staticmethod
    def _matrix_args(C):
        torch = _torch()
        if C.dim() == 2:
            C = C.unsqueeze(0)
        if C.dim() != 3 or C.shape[1] != C.shape[2]:
            raise ValueError("C must be [n,n] or [batch,n,n] (square instances)")
        if C.dtype not in (torch.float32, torch.float64):
            raise ValueError("C must be float32 (exactly representable instances) or float64")
        C = C.contiguous()
        return C, int(C.dtype == torch.float64), int(C.shape[0]), int(C.shape[1])

    @_ordered
    def narrow(self, C64):
        """float64 CUDA tensor -> (float32 copy, exact flag).  One pass on the device."""
        torch = _torch()
        C64 = C64.contiguous()
        out = torch.empty_like(C64, dtype=torch.float32)
        flag = torch.ones(1, dtype=torch.int32, device=C64.device)
        check(self.lib.b200lap_dev_narrow(self.handle, ptr(C64), C64.numel(), ptr(out), ptr(flag)), "b200lap_dev_narrow", self.lib)
        self.sync()
        return out, bool(flag.item())

    @_ordered
    def col_argmin(self, C):
        torch = _torch()
        C, f64, B, n = self._matrix_args(C)
        colmin = self._empty((B, n), C.dtype)
        colarg = self._empty((B, n), torch.int32)
        check(self.lib.b200lap_dev_col_argmin(self.handle, ptr(C), f64, B, n, ptr(colmin), ptr(colarg)), "b200lap_dev_col_argmin", self.lib)
        return colmin, colarg

    @_ordered
    def row_features(self, C, topk: int = 0):
        torch = _torch()
        C, f64, B, n = self._matrix_args(C)
        feat = self._empty((B, n, ROW_FEAT_DIM), torch.float32)
        topv = self._empty((B, n, max(topk, 1)), torch.float32)
        check(self.lib.b200lap_dev_row_features(self.handle, ptr(C), f64, B, n, int(topk), None, ptr(feat), ptr(topv) if topk > 0 else None),
              "b200lap_dev_row_features", self.lib)
        return feat, (topv if topk > 0 else None)

    @_ordered
    def onegnn_forward(self, model: "Model", feat, topv=None, want_raw: bool = False):
        torch = _torch()
        feat = feat.contiguous()
        B, n = int(feat.shape[0]), int(feat.shape[1])
        u = self._empty((B, n), torch.float32)
        raw = self._empty((B, n), torch.float32) if want_raw else None
        has_cost = int(topv is not None)
        if topv is not None:
            topv = topv.contiguous()
        check(self.lib.b200lap_dev_onegnn_forward(self.handle, model.handle, ptr(feat), ptr(topv), has_cost, B, n, ptr(u), ptr(raw)),
              "b200lap_dev_onegnn_forward", self.lib)
        return (u, raw) if want_raw else u

    @_ordered
    def min_trick(self, C, u32):
        torch = _torch()
        C, f64, B, n = self._matrix_args(C)
        u32 = u32.contiguous().view(B, n)
        v = self._empty((B, n), torch.float64)
        check(self.lib.b200lap_dev_min_trick(self.handle, ptr(C), f64, B, n, ptr(u32), ptr(v)), "b200lap_dev_min_trick", self.lib)
        return v

    @_ordered
    def predict_duals(self, model: "Model", C, want_features: bool = False):
        """features -> OneGNN -> min-trick; returns (u64, v64, u32[, feat]) CUDA tensors [B,n]."""
        torch = _torch()
        C, f64, B, n = self._matrix_args(C)
        u64 = self._empty((B, n), torch.float64)
        v64 = self._empty((B, n), torch.float64)
        u32 = self._empty((B, n), torch.float32)
        feat = self._empty((B, n, ROW_FEAT_DIM), torch.float32) if want_features else None
        check(self.lib.b200lap_dev_predict_duals(self.handle, model.handle, ptr(C), f64, B, n, ptr(u64), ptr(v64), ptr(u32), ptr(feat)),
              "b200lap_dev_predict_duals", self.lib)
        return (u64, v64, u32, feat) if want_features else (u64, v64, u32)

    @_ordered
    def solve_seeded(self, C, u, v, eps: float = 1e-12, want_trace: bool = False):
        """Seeded JV on device buffers -> (x int32[B,n], y int32[B,n], rc int32[B][, trace int64[B,12]])."""
        torch = _torch()
        C, f64, B, n = self._matrix_args(C)
        u = u.contiguous().view(B, n)
        v = v.contiguous().view(B, n)
        if u.dtype != torch.float64 or v.dtype != torch.float64:
            raise ValueError("seeds must be float64")
        x = torch.full((B, n), -1, dtype=torch.int32, device=C.device)
        y = torch.full((B, n), -1, dtype=torch.int32, device=C.device)
        rc = self._empty((B,), torch.int32)
        tr = self._empty((B, TRACE_WORDS), torch.int64) if want_trace else None
        check(self.lib.b200lap_dev_solve_seeded(self.handle, ptr(C), f64, B, n, ptr(u), ptr(v), float(eps), ptr(x), ptr(y), ptr(rc), ptr(tr), None),
              "b200lap_dev_solve_seeded", self.lib)
        return (x, y, rc, tr) if want_trace else (x, y, rc)

    @_ordered
    def solve_cold(self, C, want_trace: bool = False, want_v: bool = False):
        torch = _torch()
        C, f64, B, n = self._matrix_args(C)
        x = torch.full((B, n), -1, dtype=torch.int32, device=C.device)
        y = torch.full((B, n), -1, dtype=torch.int32, device=C.device)
        rc = self._empty((B,), torch.int32)
        tr = self._empty((B, TRACE_WORDS), torch.int64) if want_trace else None
        vout = self._empty((B, n), torch.float64) if want_v else None
        check(self.lib.b200lap_dev_solve_cold(self.handle, ptr(C), f64, B, n, ptr(x), ptr(y), ptr(rc), ptr(tr), ptr(vout)),
              "b200lap_dev_solve_cold", self.lib)
        out = [x, y, rc]
        if want_trace:
            out.append(tr)
        if want_v:
            out.append(vout)
        return tuple(out)

    @_ordered
    def bf_duals(self, C, x):
        """Column potentials of solvers/dual_computation.py:32-47 (Bellman-Ford over the difference constraints of an
        optimal matching x: row -> column) for every instance of the batch -> (v f64[B,n], rounds)."""
        import ctypes
        torch = _torch()
        C, f64, B, n = self._matrix_args(C)
        x = x.to(torch.int32).contiguous().view(B, n)
        v = self._empty((B, n), torch.float64)
        rounds = ctypes.c_int(0)
        check(self.lib.b200lap_dev_bf_duals(self.handle, ptr(C), f64, B, n, ptr(x), ptr(v), ctypes.byref(rounds)), "b200lap_dev_bf_duals", self.lib)
        return v, int(rounds.value)

    @_ordered
    def row_features_torch(self, C, topk: int = 0):
        """compute_row_features_torch's definitions (gnn/features.py:246-351) on a float32 device batch -> feat [B,n,21]."""
        self.set_option("feat_torch_mode", 1)
        try:
            return self.row_features(C, topk=topk)[0]
        finally:
            self.set_option("feat_torch_mode", 0)

    @_ordered
    def front_end(self, C, u, v, eps: float = 1e-12):
        """The solver's front-end sweep alone -> (u_tight f64[B,n], tight_cnt i32[B,n], any_violation bool[B],
        infeasible bool[B], total_tight int64[B])."""
        torch = _torch()
        C, f64, B, n = self._matrix_args(C)
        u = u.contiguous().view(B, n)
        v = v.contiguous().view(B, n)
        ut = self._empty((B, n), torch.float64)
        tc = self._empty((B, n), torch.int32)
        fl = self._empty((B, 4), torch.int32)
        check(self.lib.b200lap_dev_front_end(self.handle, ptr(C), f64, B, n, ptr(u), ptr(v), float(eps), ptr(ut), ptr(tc), ptr(fl)),
              "b200lap_dev_front_end", self.lib)
        self.sync()
        total = (fl[:, 2].to(torch.int64) & 0xFFFFFFFF) | (fl[:, 3].to(torch.int64) << 32)
        return ut, tc, fl[:, 0] != 0, fl[:, 1] != 0, total

    def pipeline(self, model: "Model", C, eps: float = 1e-12, want_trace: bool = False):
        """features -> OneGNN -> min-trick -> seeded JV without leaving the device.

        With ``set_overlap(k)`` consecutive calls rotate through k of the context's lanes (k independent batches in
        flight: one 64-instance solve occupies 64 of the 148 SMs); the returned tensors are then complete only after
        ``sync()`` -- or, for work on torch's current stream, after ``join()`` + ``wait_for_context()``."""
        torch = _torch()
        cur = torch.cuda.current_stream(self.device)
        overlap = getattr(self, "_overlap", False)
        lanes = [self.lane_stream(k) for k in range(overlap)] if overlap else [self.torch_stream()]
        for st in lanes:
            if cur.cuda_stream != st.cuda_stream:
                st.wait_stream(cur)                     # C (and the model) are complete before our kernels read them
        with torch.cuda.device(self.device):
            C, f64, B, n = self._matrix_args(C)
            x = torch.full((B, n), -1, dtype=torch.int32, device=C.device)
            y = torch.full((B, n), -1, dtype=torch.int32, device=C.device)
            rc = self._empty((B,), torch.int32)
            u64 = self._empty((B, n), torch.float64)
            v64 = self._empty((B, n), torch.float64)
            tr = self._empty((B, TRACE_WORDS), torch.int64) if want_trace else None
            if overlap:
                for st in lanes:                         # the fills above ran on torch's current stream
                    st.wait_stream(cur)
            check(self.lib.b200lap_dev_pipeline(self.handle, model.handle, ptr(C), f64, B, n, float(eps), ptr(x), ptr(y), ptr(rc), ptr(u64), ptr(v64), ptr(tr)),
                  "b200lap_dev_pipeline", self.lib)
            if overlap:
                # The tensors were allocated on torch's current stream but are written (u64 / v64 also read back) by the lane's
                # kernels, which may still be queued when the caller drops them: without this torch's caching allocator hands
                # the memory to the next call at once, and a later batch's duals can land in a block an earlier lane's solve
                # is about to write its assignment into (seen once as "the lanes disagree" and as a 34 ms step at N = 2).
                used = self.lane_stream(int(self.lib.b200lap_ctx_last_lane(self.handle)))
                for t in (x, y, rc, u64, v64, tr):
                    if t is not None:
                        t.record_stream(used)
        if not overlap and cur.cuda_stream != lanes[0].cuda_stream:
            cur.wait_stream(lanes[0])
        return (x, y, rc, u64, v64, tr) if want_trace else (x, y, rc, u64, v64)

    # ---- two batches in flight ------------------------------------------------------------------------------
    def lane_stream(self, lane: int):
        torch = _torch()
        if getattr(self, "_lanes", None) is None:
            self._lanes = [torch.cuda.ExternalStream(int(self.lib.b200lap_ctx_lane_stream(self.handle, k) or 0), device=self.device)
                           for k in range(8)]
        return self._lanes[lane]

    def set_overlap(self, lanes):
        """Rotate whole-pipeline calls through `lanes` lanes (2..8; False/0/1 = off; see ``pipeline``)."""
        lanes = 2 if lanes is True else int(lanes or 0)
        lanes = 0 if lanes < 2 else min(lanes, 8)
        self._overlap = lanes
        self.set_option("overlap_steps", lanes)

    def last_lane_event(self):
        """A CUDA event recorded behind the most recent ``pipeline`` call on the lane it was enqueued on: ``synchronize()``
        returns when that batch is done (used as the handle of ``b200lap.drain_queue``)."""
        torch = _torch()
        lane = int(self.lib.b200lap_ctx_last_lane(self.handle))
        ev = torch.cuda.Event()
        ev.record(self.lane_stream(lane) if getattr(self, "_overlap", 0) else self.torch_stream())
        return ev

    def join(self):
        """Lane 0's stream (``torch_stream()``) waits on the device for everything enqueued on lane 1."""
        check(self.lib.b200lap_ctx_join(self.handle), "b200lap_ctx_join", self.lib)


class HostPipeline:
    """Throughput form of the host-buffer pipeline (``b200lap_pipeline_batch_submit`` / ``_wait``): keep two batches
    in flight so that the upload and dense pass of one overlap the solve of the other.

        hp = HostPipeline(state_dict)
        t = hp.submit(C_pinned_f64, x_i64, y_i64, rc_i32)      # returns at once
        ...
        hp.wait(t)                                              # x, y, rc are complete
    """

    def __init__(self, state_dict: Mapping, topk: int = 16, lanes: int = 2):
        self.lib = _lib.load()
        self.ctx = self.lib.b200lap_default_ctx()
        if not self.ctx:
            raise B200LapError("no CUDA device is visible: b200lap has no CPU path")
        self.lanes = max(2, min(int(lanes), 8))
        check(self.lib.b200lap_ctx_set_option(self.ctx, b"overlap_steps", self.lanes), "b200lap_ctx_set_option", self.lib)
        blob, in_dim, hidden, layers = pack_state_dict(state_dict)
        h = ctypes.c_void_p()
        check(self.lib.b200lap_model_create(self.ctx, blob.ctypes.data, blob.size, in_dim, hidden, layers, int(topk), ctypes.byref(h)),
              "b200lap_model_create", self.lib)
        self.model = h

    def submit(self, C, x, y, rc, eps: float = 1e-12):
        B, n = int(C.shape[0]), int(C.shape[1])
        job = ctypes.c_void_p()
        check(self.lib.b200lap_pipeline_batch_submit(self.model, ptr(C), B, n, float(eps), ptr(x), ptr(y), ptr(rc), ctypes.byref(job)),
              "b200lap_pipeline_batch_submit", self.lib)
        return job

    def wait(self, job):
        check(self.lib.b200lap_pipeline_batch_wait(job), "b200lap_pipeline_batch_wait", self.lib)

    def run(self, C, x, y, rc, eps: float = 1e-12):
        """One synchronous batch (``b200lap_pipeline_batch``)."""
        B, n = int(C.shape[0]), int(C.shape[1])
        check(self.lib.b200lap_pipeline_batch(self.model, ptr(C), B, n, float(eps), ptr(x), ptr(y), ptr(rc), None, None, None),
              "b200lap_pipeline_batch", self.lib)

    def close(self):
        if getattr(self, "model", None):
            self.lib.b200lap_model_destroy(self.model)
            self.model = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Model:
    """OneGNN weights packed on a context's device (b200lap_model)."""

    def __init__(self, ctx: Context, state_dict: Mapping, topk: int = 16):
        blob, in_dim, hidden, layers = pack_state_dict(state_dict)
        self.ctx = ctx
        self.in_dim, self.hidden, self.layers, self.topk = in_dim, hidden, layers, int(topk)
        h = ctypes.c_void_p()
        check(ctx.lib.b200lap_model_create(ctx.handle, blob.ctypes.data, blob.size, in_dim, hidden, layers, int(topk), ctypes.byref(h)),
              "b200lap_model_create", ctx.lib)
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.b200lap_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: Optional[int] = None) -> Context:
    torch = _torch()
    if device is None:
        device = torch.cuda.current_device()
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def trace_dict(row) -> dict:
    return {name: int(row[k]) for k, name in enumerate(TRACE_NAMES)}
