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


def feature_close(got: np.ndarray, ref: np.ndarray, rtol: float = 1e-4, atol: float = 1e-7) -> None:
    """|got - ref| <= rtol * |ref| + atol per entry; the absolute floor covers quantities formed by
    cancellation (positional sines at multiples of pi, zero gaps)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref)
    lim = rtol * np.abs(ref) + atol
    if not np.all(err <= lim):
        bad = np.argwhere(err > lim)
        i, j = bad[0]
        raise AssertionError(f"{bad.shape[0]} feature entries off; first at row {i} feature {j}: got {got[i, j]!r} ref {ref[i, j]!r}")


def state_dict_from_golden(g, prefix: str = "small_sd/") -> dict:
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}
