"""``OneGNN`` -- host-side mirror of /root/reference/gnn/one_gnn.py:39-160.

The class keeps the reference's constructor, parameter names (so ``state_dict`` files interchange,
SURVEY.md App. C.1) and ``forward(row_feat, cost=, mask=) -> {'u': ...}`` contract, but the forward is
inference-only and runs on the fused sm_100a kernels (csrc/features.cuh top-k selection + csrc/mlp.cuh):
parameters are packed onto the device once and re-packed when they change.  Training is out of scope.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn


class ResidualBlock(nn.Module):
    """Parameter container for one residual MLP block (fc1, fc2, norm) -- one_gnn.py:18-36."""

    def __init__(self, hidden: int, dropout: float) -> None:
        super().__init__()
        self.fc1 = nn.Linear(hidden, hidden)
        self.fc2 = nn.Linear(hidden, hidden)
        self.norm = nn.LayerNorm(hidden)
        self.dropout = nn.Dropout(dropout)
        self.act = nn.GELU()


class OneGNN(nn.Module):
    def __init__(self, in_dim: int, hidden: int = 64, layers: int = 2, dropout: float = 0.1, topk: int = 16) -> None:
        super().__init__()
        if layers < 1:
            raise ValueError("layers must be >= 1")
        if hidden < 2:
            raise ValueError("hidden dimension must be >= 2 for head projection")
        self.input_proj = nn.Sequential(nn.Linear(in_dim, hidden), nn.GELU(), nn.LayerNorm(hidden))
        self.blocks = nn.ModuleList([ResidualBlock(hidden, dropout) for _ in range(layers)])
        head_hidden = max(hidden // 2, 1)
        self.pre_out = nn.Linear(hidden, 1)
        self.row_out = nn.Sequential(nn.Linear(hidden, head_hidden), nn.GELU(), nn.Dropout(dropout), nn.Linear(head_hidden, 1))
        self.topk = topk
        self.edge_mlp = nn.Sequential(nn.Linear(1, hidden), nn.GELU(), nn.Linear(hidden, hidden))
        self.message_norm = nn.LayerNorm(hidden)
        self.message_dropout = nn.Dropout(dropout)
        self._packed = None
        self._packed_key = None

    # ---- device plumbing ---------------------------------------------------------------------------
    def _device_model(self, ctx):
        key = (ctx.device, tuple(int(p._version) for p in self.parameters()), tuple(p.data_ptr() for p in self.parameters()))
        if self._packed is None or self._packed_key != key:
            from b200lap.runtime import Model
            self._packed = Model(ctx, self.state_dict(), topk=self.topk)
            self._packed_key = key
        return self._packed

    def forward(self, row_feat: torch.Tensor, *, cost: Optional[torch.Tensor] = None,
                mask: Optional[torch.Tensor] = None) -> dict:
        if self.training:
            raise RuntimeError("the B200 OneGNN is inference-only: call .eval() first (training is out of scope)")
        if row_feat.ndim == 2:
            row_feat = row_feat.unsqueeze(0)
        if row_feat.ndim != 3:
            raise ValueError("row_feat must have shape (batch, n, F)")
        from b200lap.runtime import default_context
        dev_index = row_feat.device.index if row_feat.is_cuda else None
        ctx = default_context(dev_index)
        model = self._device_model(ctx)
        dev = torch.device(f"cuda:{ctx.device}")
        B, n, _ = row_feat.shape
        feat = row_feat.to(dev, torch.float32).contiguous()
        if mask is not None and mask.ndim == 1:
            mask = mask.unsqueeze(0)
        topv = None
        if cost is not None:
            if cost.ndim == 2:
                cost = cost.unsqueeze(0)
            k = min(self.topk, cost.size(-1))
            if k > 0 and n > 0:
                c = cost.to(dev)
                if c.dtype not in (torch.float32, torch.float64):
                    c = c.float()
                _, topv = ctx.row_features(c, topk=self.topk)
                if mask is not None:
                    # masked rows see +inf everywhere (one_gnn.py:144-145): no valid edge, zero message
                    topv = topv.masked_fill(~mask.to(dev).unsqueeze(-1), float("inf"))
        u = ctx.onegnn_forward(model, feat, topv)
        if mask is not None:
            u = u.masked_fill(~mask.to(dev), 0.0)
        return {"u": u.to(row_feat.device) if not row_feat.is_cuda else u}
