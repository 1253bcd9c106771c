"""CPU tests: pin the NumPy restatements of the dense half (oracle/features_np.py,
oracle/onegnn_np.py, oracle/pipeline_np.py) against outputs of the reference's own Python
recorded in tests/golden/dense_golden.npz (see tests/golden/make_dense_golden.py)."""
import os

import numpy as np
import pytest

import oracle
from oracle import features_np, onegnn_np, pipeline_np
from _fixtures import GOLDEN, feature_close

FAMILIES = ("uniform", "sparse", "sparse1e6", "metric", "clustered", "odd", "tiny")


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "dense_golden.npz"))


def small_sd(g):
    return {k[len("small_sd/"):]: g[k] for k in g.files if k.startswith("small_sd/")}


@pytest.mark.parametrize("fam", FAMILIES)
def test_features_match_reference(g, fam):
    C = g[f"{fam}/C"]
    f = features_np.row_features(C)
    assert f.dtype == np.float32 and f.shape == (C.shape[0], 21)
    feature_close(f, g[f"{fam}/feat"], rtol=1e-6)


@pytest.mark.parametrize("fam", FAMILIES)
def test_small_model_matches_reference(g, fam):
    C = g[f"{fam}/C"]
    sd = small_sd(g)
    feat = g[f"{fam}/feat"]
    mask = np.ones(C.shape[0], dtype=bool)
    u, raw = onegnn_np.forward(sd, feat, cost=C.astype(np.float32), mask=mask, topk=8, return_raw=True)
    scale = np.abs(g[f"{fam}/small_raw"]).max()
    assert np.abs(raw - g[f"{fam}/small_raw"]).max() <= 2e-5 * scale
    assert np.abs(u - g[f"{fam}/small_u"]).max() <= 2e-5 * scale
    # v is exact GIVEN u: feed the reference's u
    assert np.array_equal(pipeline_np.min_trick(C, g[f"{fam}/small_u"]), g[f"{fam}/small_v"])


def test_named_model_rebuilds_and_matches(g):
    torch = pytest.importorskip("torch")
    from gnn.one_gnn import OneGNN  # the repo's host-side mirror (CPU construction only)
    torch.manual_seed(0)
    model = OneGNN(21, hidden=192, layers=4, dropout=0.1, topk=16).eval()
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    for k, a in sd.items():
        a = a.astype(np.float64)
        ck = np.array([a.sum(), np.abs(a).sum(), (a * np.arange(1, a.size + 1).reshape(a.shape)).sum()])
        assert np.array_equal(ck, g["named_ck/" + k]), k
    assert sum(a.size for a in sd.values()) == 359234
    for fam in ("uniform", "sparse", "sparse1e6", "metric", "clustered"):
        C = g[f"{fam}/C"]
        mask = np.ones(C.shape[0], dtype=bool)
        u, raw = onegnn_np.forward(sd, g[f"{fam}/feat"], cost=C.astype(np.float32), mask=mask, return_raw=True)
        scale = np.abs(g[f"{fam}/named_raw"]).max()
        assert np.abs(u - g[f"{fam}/named_u"]).max() <= 2e-5 * scale, fam
        assert np.array_equal(pipeline_np.min_trick(C, g[f"{fam}/named_u"]), g[f"{fam}/named_v"])


def test_pipeline_solves_to_optimum(g):
    sd = small_sd(g)
    for fam in ("uniform", "metric", "clustered"):
        C = g[f"{fam}/C"]
        x, y, cost, u, v = pipeline_np.solve(C, sd, topk=8)
        xc, yc = oracle.port_lapjv_internal(C)
        assert cost == pytest.approx(C[np.arange(C.shape[0]), xc].sum(), rel=1e-12)
        assert np.array_equal(np.sort(x), np.arange(C.shape[0]))
        assert (C - u[:, None] - v[None, :]).min() >= -1e-9   # min-trick duals are feasible


def test_front_end_numpy_statements():
    rng = np.random.default_rng(3)
    C = rng.uniform(0, 1, (40, 40))
    u, v = pipeline_np.project_feasible(C, rng.normal(0.3, 0.2, 40), rng.normal(0.3, 0.2, 40))
    assert (C - u[:, None] - v[None, :]).min() >= -1e-12
    R = pipeline_np.reduce_costs(C, u, v)
    assert R.min() >= 0 and R.flags.c_contiguous
    # tight count of the numpy statement == the C port's counter on already-feasible seeds
    v0 = np.min(C, axis=0)
    rc, uu, vv, x, y, tr = oracle.port_front_end(C, np.zeros(40), v0)
    assert rc == 0 and tr["tight_edges"] == pipeline_np.tight_edge_count(C, uu, vv)
