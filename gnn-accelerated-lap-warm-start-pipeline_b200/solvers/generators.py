"""Synthetic cost-matrix families used as inputs of the hot path (not accelerated).

Same laws, seeds and function names as the reference generators, restated in
vectorised NumPy (the reference's metric generator is an O(n^2) Python loop):

  generate_uniform_costs     /root/reference/solvers/generators.py:12-24
  generate_sparse_costs      /root/reference/solvers/generators.py:60-94   (fill 100.0)
  generate_metric_costs      /root/reference/solvers/generators.py:97-110
  generate_clustered_costs   /root/reference/solvers/generators.py:113-123
  generate_sparse_dataset_costs  /root/reference/data/generators.py:56-69  (fill 1e6)

``snap_to_fp32_grid`` rounds an instance to values that binary32 holds exactly,
so the host fp64 matrix and the device fp32 copy are the same real numbers
(SURVEY.md 8d); every benchmark/parity instance goes through it.
"""
from __future__ import annotations

import numpy as np

FAMILIES = ("uniform", "sparse", "metric", "clustered")


def snap_to_fp32_grid(C: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(C, dtype=np.float64).astype(np.float32).astype(np.float64)


def generate_uniform_costs(n: int, seed: int = 42) -> np.ndarray:
    rs = np.random.RandomState(seed)
    return rs.uniform(0.0, 1.0, (n, n)).astype(np.float64)


def _ensure_cover(keep: np.ndarray, draw) -> None:
    """Every row, then every column, keeps at least one edge (same draw order as the reference)."""
    n = keep.shape[0]
    for i in np.flatnonzero(~keep.any(axis=1)):
        keep[i, draw(n)] = True
    for j in np.flatnonzero(~keep.any(axis=0)):
        keep[draw(n), j] = True


def generate_sparse_costs(n: int, sparsity_ratio: float = 0.3, seed: int = 42) -> np.ndarray:
    rs = np.random.RandomState(seed)
    C = rs.uniform(0.1, 1.0, (n, n))
    keep = rs.random_sample((n, n)) < sparsity_ratio
    _ensure_cover(keep, lambda m: rs.randint(m))
    return np.where(keep, C, 100.0).astype(np.float64)


def generate_metric_costs(n: int, seed: int = 42) -> np.ndarray:
    rs = np.random.RandomState(seed)
    pts = rs.uniform(0, 100, (n, 2))
    d = pts[:, None, :] - pts[None, :, :]
    # np.linalg.norm of a 2-vector is sqrt(dx*dx + dy*dy)
    return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])


def generate_clustered_costs(n: int, blocks: int = 4, noise: float = 0.1, seed: int = 42) -> np.ndarray:
    rng = np.random.default_rng(seed)
    C = rng.uniform(0.0, 1.0, size=(n, n))
    size = max(1, n // max(1, blocks))
    for b in range(blocks):
        lo = b * size
        hi = n if b == blocks - 1 else min(n, (b + 1) * size)
        C[lo:hi, lo:hi] -= 0.4
    C += noise * rng.normal(0.0, 1.0, size=(n, n))
    return np.maximum(C, 0.0).astype(np.float64)


def generate_sparse_dataset_costs(n: int, sparsity: float = 0.3, seed: int = 42) -> np.ndarray:
    """The dataset flavour of "sparse": U(0,1) costs, forbidden edges = 1e6 (collides with LARGE)."""
    rng = np.random.default_rng(seed)
    dense = generate_uniform_costs(n, seed=int(rng.integers(0, np.iinfo(np.uint32).max)))
    keep = rng.random(size=(n, n)) < sparsity
    _ensure_cover(keep, lambda m: rng.integers(0, m))
    dense[~keep] = 1e6
    return dense.astype(np.float64)


_BY_NAME = {
    "uniform": generate_uniform_costs,
    "sparse": generate_sparse_costs,
    "sparse1e6": generate_sparse_dataset_costs,
    "metric": generate_metric_costs,
    "clustered": generate_clustered_costs,
}


def make_instance(family: str, n: int, seed: int = 42, snap: bool = True) -> np.ndarray:
    C = _BY_NAME[family](n, seed=seed)
    return snap_to_fp32_grid(C) if snap else np.ascontiguousarray(C, dtype=np.float64)


def mixed_batch(n: int, batch: int, first_seed: int = 42, families=FAMILIES):
    """``batch`` instances cycling through ``families``; instance ``k`` uses seed ``first_seed + k``
    (scripts/gnn_large_scale_benchmark.py:415 seeds instances the same way)."""
    return [(families[k % len(families)], make_instance(families[k % len(families)], n, first_seed + k))
            for k in range(batch)]
