"""GNNPredictor -- the reference's inference glue, on the device.

Mirrors ``GNNPredictor.predict(C) -> (u64, v64)`` of /root/reference/scripts/gnn_benchmark.py:213-289
(features -> OneGNN with the cost branch and an all-true mask -> min-trick -> float64 casts), with
the same dtype path as its CPU branch: statistics of the binary64 matrix, model in binary32, v in binary64.
"""
from __future__ import annotations

from typing import Mapping, Optional, Union

import numpy as np

from .runtime import Context, Model, default_context, _torch


def _load_state_dict(source) -> tuple:
    """Accepts a OneGNN module, a state_dict, or a checkpoint path in the reference's format
    (/root/reference/gnn/train_one_gnn.py:409-420: {'model_state_dict', 'hidden_dim', 'layers', ...})."""
    topk = None
    if hasattr(source, "state_dict") and callable(source.state_dict):
        topk = getattr(source, "topk", None)
        return source.state_dict(), topk
    if isinstance(source, (str, bytes)):
        torch = _torch()
        ck = torch.load(source, map_location="cpu", weights_only=False)
        if isinstance(ck, Mapping) and "model_state_dict" in ck:
            cfg = ck.get("config", {}) if isinstance(ck.get("config", {}), Mapping) else {}
            topk = ck.get("topk", cfg.get("topk"))
            return ck["model_state_dict"], topk
        return ck, topk
    return source, topk


class GNNPredictor:
    def __init__(self, model_source, device: Union[int, str, None] = None, topk: Optional[int] = None,
                 ctx: Optional[Context] = None):
        torch = _torch()
        if isinstance(device, str):
            device = torch.device(device).index or 0
        self.ctx = ctx or default_context(device)
        sd, ck_topk = _load_state_dict(model_source)
        self.model = Model(self.ctx, sd, topk=topk if topk is not None else (ck_topk if ck_topk is not None else 16))

    def to_device(self, C):
        """host float64 matrix (or batch) -> device tensor, binary32 when that is exact."""
        torch = _torch()
        if isinstance(C, np.ndarray):
            C = torch.from_numpy(np.ascontiguousarray(C, dtype=np.float64))
        C = C.to(f"cuda:{self.ctx.device}", non_blocking=False)
        if C.dtype == torch.float64:
            C32, exact = self.ctx.narrow(C)
            return C32 if exact else C
        return C

    def predict_device(self, C_dev):
        """device matrix/batch -> (u64, v64) CUDA tensors."""
        u64, v64, _ = self.ctx.predict_duals(self.model, C_dev)
        return u64, v64

    def predict(self, C):
        """numpy matrix -> (u float64[n], v float64[n]) numpy, like the reference."""
        Cd = self.to_device(C)
        u64, v64 = self.predict_device(Cd)
        self.ctx.sync()
        u = u64.cpu().numpy()
        v = v64.cpu().numpy()
        if np.asarray(C).ndim == 2:
            return u[0], v[0]
        return u, v
