"""CPU tests: pin the C restatement (oracle/jv_port.c) against the unmodified reference
solver (oracle/_ref) and against the reference's own known-answer vectors."""
import numpy as np
import pytest

import oracle
from solvers import generators as gen
from _fixtures import (SEEDED_INT_FIXTURES, load_known_answers, mintrick_seeds, noisy_oracle_seeds,
                       seeded_int_case)

needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (no /root/reference)")


def _same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]


def test_known_answers_small():
    g = load_known_answers()
    for k in range(int(g["n_small"])):
        C = g[f"small{k}_C"]
        x, y = oracle.port_lapjv_internal(C)
        assert list(x) == list(g[f"small{k}_x"]), k
        assert list(y) == list(g[f"small{k}_y"]), k
        assert C[np.arange(C.shape[0]), x].sum() == float(g[f"small{k}_opt"])


@pytest.mark.parametrize("name", sorted(SEEDED_INT_FIXTURES))
def test_known_answers_seeded_int(name):
    g = load_known_answers()
    C, opt, xg, yg = seeded_int_case(g, name)
    x, y = oracle.port_lapjv_internal(C)
    assert C[np.arange(C.shape[0]), x].sum() == opt
    assert np.array_equal(x, xg) and np.array_equal(y, yg)


def test_known_answer_eps_and_arr_loop():
    g = load_known_answers()
    C = g["eps_C"]
    x, y = oracle.port_lapjv_internal(C)
    assert C[np.arange(C.shape[0]), x].sum() == pytest.approx(float(g["eps_opt"]), rel=1e-13)
    assert np.array_equal(x, g["eps_x"]) and np.array_equal(y, g["eps_y"])
    opt, ind1, ind0 = oracle.lapjv_py(g["arr_C"], extend_cost=True, return_cost=True)
    assert opt == pytest.approx(float(g["arr_opt"]), rel=1e-10)
    assert list(ind0) in ([5, 1, 2], [1, 5, 2])


def test_seeded_prints_of_reference_scripts():
    # LAP/test_seeded.py:8-12 (3x3, zero seeds) and LAP/demo_seeded.py:18-37 (4x4)
    C = np.array([[4.0, 1.0, 3.0], [2.0, 0.0, 5.0], [3.0, 2.0, 2.0]])
    x, y, cost = oracle.port_lapjv_seeded(C, np.zeros(3), np.zeros(3))
    assert cost == 5.0 and list(x) == [1, 0, 2]
    C = np.array([[4.0, 2.0, 8.0, 6.0], [6.0, 4.0, 1.0, 2.0], [8.0, 6.0, 4.0, 3.0], [2.0, 8.0, 5.0, 7.0]])
    x, y, cost = oracle.port_lapjv_seeded(C, np.zeros(4), np.array([2.0, 2.0, 1.0, 2.0]))
    assert cost == 8.0 and list(x) == [1, 2, 3, 0] and list(y) == [3, 0, 1, 2]
    # the "infeasible" seeds of the demo do NOT raise in the reference: projection repairs them
    x, y, cost = oracle.port_lapjv_seeded(C, np.full(4, 10.0), np.zeros(4))
    assert cost == 8.0


def test_argument_errors():
    C = np.zeros((3, 3))
    with pytest.raises(ValueError):
        oracle.port_lapjv_seeded(C, np.zeros(2), np.zeros(3))
    with pytest.raises(RuntimeError):
        oracle.port_lapjv_seeded(np.zeros((2, 3)), np.zeros(2), np.zeros(3))  # rc -4


@needs_ref
@pytest.mark.parametrize("family", ["uniform", "sparse", "sparse1e6", "metric", "clustered"])
@pytest.mark.parametrize("n", [37, 300])
def test_port_matches_reference_seeded(family, n):
    rng = np.random.default_rng(n)
    C = gen.make_instance(family, n, 42)
    seeds = [mintrick_seeds(C, rng), (np.zeros(n), np.zeros(n)),
             (rng.normal(0.2, 0.2, n), rng.normal(0.2, 0.2, n))]
    for s in (0.0, 1e-3, 1e-2):
        seeds.append(noisy_oracle_seeds(C, s))
    for u, v in seeds:
        tr = oracle.Trace()
        assert _same(oracle.ref_lapjv_seeded(C, u, v), oracle.port_lapjv_seeded(C, u, v, trace=tr))
        assert tr.rc == 0


@needs_ref
def test_port_matches_reference_ties():
    rng = np.random.default_rng(7)
    for n in (2, 3, 5, 17, 64, 150):
        for _ in range(6):
            C = rng.integers(0, 4, (n, n)).astype(np.float64)
            xr, yr = oracle.ref_lapjv_internal(C)
            xp, yp = oracle.port_lapjv_internal(C)
            assert np.array_equal(xr, xp) and np.array_equal(yr, yp)
            for u, v in ((np.zeros(n), C.min(axis=0)),
                         (rng.integers(0, 3, n).astype(float), rng.integers(0, 3, n).astype(float))):
                assert _same(oracle.ref_lapjv_seeded(C, u, v), oracle.port_lapjv_seeded(C, u, v))


@needs_ref
def test_port_matches_reference_large_sentinel():
    # values straddling LARGE=1e6 exercise the column-reduction / ARR sentinel rules
    rng = np.random.default_rng(11)
    for n in (8, 40, 120):
        C = rng.uniform(0, 1, (n, n))
        C[rng.random((n, n)) < 0.6] = 1e6
        C[np.arange(n), rng.permutation(n)] = rng.uniform(0, 1, n)
        C[:, 0] = 1e6
        C[rng.integers(0, n), 0] = 0.5
        xr, yr = oracle.ref_lapjv_internal(C)
        xp, yp = oracle.port_lapjv_internal(C)
        assert np.array_equal(xr, xp) and np.array_equal(yr, yp)
        C2 = C * 3.0
        xr, yr = oracle.ref_lapjv_internal(C2)
        xp, yp = oracle.port_lapjv_internal(C2)
        assert np.array_equal(xr, xp) and np.array_equal(yr, yp)


def test_oracle_seed_phases():
    """sigma=0 oracle seeds stay on the warm path; sigma>0 seeds project and fall back."""
    C = gen.make_instance("uniform", 200, 42)
    tr = oracle.Trace()
    u, v = noisy_oracle_seeds(C, 0.0)
    oracle.port_lapjv_seeded(C, u, v, trace=tr)
    assert tr.took_fallback == 0 and tr.proj_triggers == 0
    u, v = noisy_oracle_seeds(C, 1e-2)
    oracle.port_lapjv_seeded(C, u, v, trace=tr)
    assert tr.took_fallback == 1 and tr.proj_triggers > 0 and tr.tight_edges == 200
