"""Shared instance builders for the tests (restating the reference's own test helpers)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dense_int(sz: int, rng: int, hard: bool = True, seed: int = 1299821) -> np.ndarray:
    """/root/reference/LAP/lap/tests/test_utils.py:7-33 (get_dense_int + make_hard)."""
    rs = np.random.RandomState(seed)
    cost = rs.randint(1, rng + 1, size=(sz, sz))
    if hard:
        cost = cost.copy()
        for row in range(sz):
            cost[row, :] += rs.randint(0, rng)
        for col in range(sz):
            cost[:, col] += rs.randint(0, rng)
    return cost


SEEDED_INT_FIXTURES = {
    "d100": (100, 100, False),
    "d100h": (100, 100, True),
    "d1k": (1000, 100, False),
    "d1kh": (1000, 100, True),
}


def load_known_answers():
    return np.load(os.path.join(GOLDEN, "lapjv_known_answers.npz"))


def seeded_int_case(g, name):
    sz, rng, hard = SEEDED_INT_FIXTURES[name]
    c = dense_int(sz, rng, hard).astype(np.float64)
    assert float(c.sum()) == float(g[f"{name}_sum"])
    assert float((c * np.arange(1, sz + 1)[None, :]).sum()) == float(g[f"{name}_wsum"])
    return c, float(g[f"{name}_opt"]), g[f"{name}_x"], g[f"{name}_y"]


def mintrick_seeds(C: np.ndarray, rng: np.random.Generator, scale: float = 0.01):
    """(u fp32-valued, v = min_i(C - u)) -- the shape of seeds the dense half produces."""
    n = C.shape[0]
    u = rng.normal(0.0, scale, n).astype(np.float32).astype(np.float64)
    v = np.min(C - u[:, None], axis=0)
    return u, v


def noisy_oracle_seeds(C: np.ndarray, sigma: float, seed: int = 42):
    """Optimal duals (from a cold JV run of the C port) + N(0, sigma), no re-projection --
    the seed law of /root/reference/solvers/dual_computation.py:77-115 (np.random.seed(42))."""
    import oracle
    _, _, u, v = oracle.port_optimal_duals(C)
    if sigma > 0:
        rs = np.random.RandomState(seed)
        u = u + rs.normal(0, sigma, u.shape[0])
        v = v + rs.normal(0, sigma, v.shape[0])
    return u, v
