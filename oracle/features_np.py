"""oracle/features_np.py -- TEST INFRASTRUCTURE ONLY (the checker, never the product).

NumPy restatement of the reference's 21-D row features,
``/root/reference/gnn/features.py:161-243`` (compute_row_features) and ``:21-31``
(_positional_encodings).  All statistics are evaluated in binary64 and cast to
binary32 at the very end, like the reference.

Parity status: PINNED -- tests/test_oracle_dense.py checks this file against
tests/golden/dense_golden.npz, which tests/golden/make_dense_golden.py produced by
importing the reference's own ``gnn.features.compute_row_features`` in the authoring
container.

Column order (features.py:223-241):
  0 row_min  1 row_max  2 row_mean  3 row_std(ddof 0)  4 row_mad  5 row_entropy
  6 second_best_gap  7 competition  8 k_mean  9 k_std  10 difficulty
  11 near_best  12 is_col_best  13..20 sin/cos positional terms
"""
from __future__ import annotations

import numpy as np

POS_FREQS = (1, 2, 4, 8)   # features.py:16
EPS = 1e-9                  # features.py:18
ROW_FEAT_DIM = 13 + 2 * len(POS_FREQS)


def positional_terms(n: int) -> np.ndarray:
    """features.py:21-31: sin/cos of 2*pi*i*f/max(1, n-1) for f in POS_FREQS -> f32[n, 8]."""
    out = np.zeros((max(n, 0), 2 * len(POS_FREQS)), dtype=np.float64)
    if n <= 0:
        return out.astype(np.float32)
    idx = np.arange(n, dtype=np.float64)
    denom = max(1, n - 1)
    for c, f in enumerate(POS_FREQS):
        ang = 2.0 * np.pi * idx * f / denom
        out[:, 2 * c] = np.sin(ang)
        out[:, 2 * c + 1] = np.cos(ang)
    return out.astype(np.float32)


def row_features(C) -> np.ndarray:
    """features.py:161-243 -> f32[n, 21]."""
    C = np.asarray(C, dtype=np.float64)
    n = C.shape[0]
    if n == 0:
        return np.zeros((0, 0), dtype=np.float32)
    m = C.shape[1]
    S = np.sort(C, axis=1)                      # ascending order statistics of every row
    lo, hi = S[:, 0], S[:, -1]

    mean = C.mean(axis=1)                       # :172
    std = C.std(axis=1)                         # :173 (ddof=0)
    med = np.median(C, axis=1)                  # :174
    mad = np.median(np.abs(C - med[:, None]), axis=1)   # :175
    mad = np.where(mad < EPS, EPS, mad)         # :176

    e = np.exp(-(C - lo[:, None]))              # :179-180
    p = e / (e.sum(axis=1, keepdims=True) + EPS)
    entropy = -(p * np.log(p + EPS)).sum(axis=1)    # :182

    if m >= 2:
        gap = S[:, 1] - S[:, 0]                 # :185-187
        competition = gap / ((hi - lo) + EPS)   # :189-191
        difficulty = 1.0 / (np.diff(S, axis=1).mean(axis=1) + EPS)   # :207-210
    else:
        gap = np.zeros(n)
        competition = np.zeros(n)
        difficulty = np.zeros(n)

    k = min(10, m)                              # :197
    kmean = S[:, :k].mean(axis=1)
    kstd = S[:, :k].std(axis=1)

    near_best = (C <= lo[:, None] * 1.1).sum(axis=1) / max(1, m)     # :215
    col_best = (C == C.min(axis=0)).sum(axis=1) / max(1, m)          # :218-219

    stats = np.stack([lo, hi, mean, std, mad, entropy, gap, competition, kmean, kstd,
                      difficulty, near_best, col_best], axis=1)
    return np.concatenate([stats.astype(np.float32), positional_terms(n)], axis=1)
