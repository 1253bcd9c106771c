"""GPU parity tests (run on the B200 with `-m gpu`): the nvcc-built libb200lap.so, called through the C ABI
and its Python mirrors, against the oracle on the same seeded inputs.

Bars: assignments/costs bit-exact against the reference lapjv_seeded semantics (oracle/jv_port.c, itself
pinned to oracle/_ref and the reference's known answers); v exact given u; features within rtol 1e-4
(+1e-7 abs) of the NumPy definition; |u - u_ref| <= 1e-4 * max|u_raw| (SURVEY.md hard-part 6).
Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

import oracle
from oracle import features_np, onegnn_np, pipeline_np
from solvers import generators as gen
from _fixtures import (GOLDEN, SEEDED_INT_FIXTURES, dense_int, feature_close, load_known_answers, mintrick_seeds,
                       noisy_oracle_seeds, seeded_int_case, state_dict_from_golden)

pytestmark = pytest.mark.gpu

FAMILIES = ("uniform", "sparse", "sparse1e6", "metric", "clustered")


@pytest.fixture(scope="module")
def ctx():
    import b200lap
    return b200lap.default_context(0)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "dense_golden.npz"))


def _oracle_seeded(C, u, v):
    tr = oracle.Trace()
    try:
        x, y, cost = oracle.port_lapjv_seeded(C, u, v, trace=tr)
        return 0, x, y, cost, tr.as_dict()
    except ValueError:
        return -3, None, None, None, tr.as_dict()


def _same_as_oracle(C, u, v, tag):
    import lap
    rco, xo, yo, co, _ = _oracle_seeded(C, u, v)
    if rco == -3:
        with pytest.raises(ValueError):
            lap.lapjv_seeded(C, u, v)
        return
    x, y, cost = lap.lapjv_seeded(C, u, v)
    assert x.dtype == np.int64 and y.dtype == np.int64
    assert np.array_equal(x, xo) and np.array_equal(y, yo) and cost == co, tag


# ---- solver half ------------------------------------------------------------------------------------
@pytest.mark.parametrize("fam", FAMILIES)
@pytest.mark.parametrize("n", [1, 2, 3, 17, 64, 300, 512, 1000])
def test_lapjv_seeded_mintrick_seeds(fam, n):
    rng = np.random.default_rng(n)
    C = gen.make_instance(fam, n, seed=42 + n)
    u, v = mintrick_seeds(C, rng)
    _same_as_oracle(C, u, v, (fam, n))


@pytest.mark.parametrize("fam", FAMILIES)
@pytest.mark.parametrize("sigma", [0.0, 1e-3, 1e-2, 1e-1])
def test_lapjv_seeded_noisy_oracle_seeds(fam, sigma):
    for n in (96, 400):
        C = gen.make_instance(fam, n, seed=7)
        u, v = noisy_oracle_seeds(C, sigma)
        _same_as_oracle(C, u, v, (fam, n, sigma))


def test_lapjv_seeded_reference_script_cases():
    import lap
    C = np.array([[4.0, 1.0, 3.0], [2.0, 0.0, 5.0], [3.0, 2.0, 2.0]])
    x, y, cost = lap.lapjv_seeded(C, np.zeros(3), np.zeros(3))
    assert cost == 5.0 and list(x) == [1, 0, 2]
    C = np.array([[4.0, 2.0, 8.0, 6.0], [6.0, 4.0, 1.0, 2.0], [8.0, 6.0, 4.0, 3.0], [2.0, 8.0, 5.0, 7.0]])
    x, y, cost = lap.lapjv_seeded(C, np.zeros(4), np.array([2.0, 2.0, 1.0, 2.0]))
    assert cost == 8.0 and list(x) == [1, 2, 3, 0] and list(y) == [3, 0, 1, 2]
    x, y, cost = lap.lapjv_seeded(C, np.full(4, 10.0), np.zeros(4))
    assert cost == 8.0
    from solvers.lap_solver import SeededLAPSolver, LAPSolver
    rows, cols, cost = SeededLAPSolver().solve(C, np.zeros(4), np.array([2.0, 2.0, 1.0, 2.0]))
    assert list(rows) == [1, 2, 3, 0] and list(cols) == [3, 0, 1, 2] and cost == 8.0
    rows, cols, cost = LAPSolver().solve(C)
    assert list(rows) == [0, 1, 2, 3] and cost == 8.0


def test_tie_heavy_integer_matrices():
    rng = np.random.default_rng(3)
    for sz, hard in ((50, False), (120, True), (257, True)):
        C = dense_int(sz, 10, hard, seed=5 + sz).astype(np.float64)
        _same_as_oracle(C, np.zeros(sz), np.zeros(sz), ("zero", sz))
        _same_as_oracle(C, np.zeros(sz), C.min(axis=0), ("colmin", sz))
        u, v = noisy_oracle_seeds(C, 0.0)
        _same_as_oracle(C, u, v, ("oracle", sz))
        _same_as_oracle(C, u + 0.5, v + 0.5, ("shifted", sz))
        _same_as_oracle(C, u + rng.normal(0, 1, sz), v, ("noisy-u", sz))


def test_binary64_storage_path():
    rng = np.random.default_rng(8)
    for n in (33, 500):
        C = rng.uniform(0, 1, (n, n))     # not representable in binary32
        u, v = mintrick_seeds(C, rng)
        _same_as_oracle(C, u, v, ("f64", n))


def test_cold_lapjv_known_answers():
    import lap
    g = load_known_answers()
    for k in range(int(g["n_small"])):
        C = g[f"small{k}_C"]
        if not np.isfinite(C).all():
            continue
        opt, x, y = lap.lapjv(C)
        assert list(x) == list(g[f"small{k}_x"]) and list(y) == list(g[f"small{k}_y"]) and opt == float(g[f"small{k}_opt"])
    for name in sorted(SEEDED_INT_FIXTURES):
        C, opt, xg, yg = seeded_int_case(g, name)
        o, x, y = lap.lapjv(C)
        assert o == opt and np.array_equal(x, xg) and np.array_equal(y, yg), name
    C = g["eps_C"]
    o, x, y = lap.lapjv(C)
    assert np.array_equal(x, g["eps_x"]) and np.array_equal(y, g["eps_y"])
    opt, ind1, ind0 = lap.lapjv(g["arr_C"], extend_cost=True, return_cost=True)
    assert opt == pytest.approx(float(g["arr_opt"]), rel=1e-10)
    # cost_limit padding path of the binding
    C = gen.make_instance("uniform", 20, 3)[:, :15]
    got = lap.lapjv(C, extend_cost=True, cost_limit=0.3)
    ref = oracle.lapjv_py(C, extend_cost=True, cost_limit=0.3)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


def test_reference_binary_agrees_when_present():
    """oracle/_ref (the unmodified reference solver, prebuilt) on a few cases, when it travelled to the box."""
    if not os.path.exists(os.path.join(os.path.dirname(oracle.__file__), "_ref", "libreflap.so")):
        pytest.skip("oracle/_ref not present")
    import lap
    rng = np.random.default_rng(21)
    for fam in FAMILIES:
        C = gen.make_instance(fam, 256, seed=11)
        u, v = mintrick_seeds(C, rng)
        xr, yr, cr = oracle.ref_lapjv_seeded(C, u, v)
        x, y, c = lap.lapjv_seeded(C, u, v)
        assert np.array_equal(x, xr) and np.array_equal(y, yr) and c == cr, fam


def test_batched_device_solve_mid2048(ctx):
    """config 2 shape at reduced batch: mixed-family n=2048, device-resident, against the oracle."""
    import torch
    n, B = 2048, 8
    rng = np.random.default_rng(1)
    batch = gen.mixed_batch(n, B, first_seed=42)
    Cs = np.stack([c for _, c in batch])
    us, vs = zip(*[mintrick_seeds(c, rng) for c in Cs])
    Cd = torch.from_numpy(Cs.astype(np.float32)).cuda()
    x, y, rc, tr = ctx.solve_seeded(Cd, torch.from_numpy(np.stack(us)).cuda(), torch.from_numpy(np.stack(vs)).cuda(), want_trace=True)
    ctx.sync()
    assert (rc == 0).all()
    x = x.cpu().numpy(); y = y.cpu().numpy(); tr = tr.cpu().numpy()
    for b in range(B):
        rco, xo, yo, co, tro = _oracle_seeded(Cs[b], us[b], vs[b])
        assert np.array_equal(x[b], xo) and np.array_equal(y[b], yo), batch[b][0]
        assert tro["tight_edges"] == tr[b][1] and tro["aug_paths"] == tr[b][7] and tro["took_fallback"] == tr[b][3]


@pytest.mark.parametrize("cluster", [2, 4, 8])
def test_cluster_mode_matches_oracle(ctx, cluster):
    """Cluster mode of the solver (relax steps spread over a thread-block cluster, state in the global workspace)
    forced on small instances: assignments and the phase counters must equal the oracle's, seeded and cold."""
    import torch
    rng = np.random.default_rng(cluster)
    cases = [(fam, n) for fam in ("uniform", "sparse", "clustered", "metric") for n in (64, 300, 1000)]
    cases.append(("ties", 256))
    try:
        ctx.set_option("solver_cluster", cluster)
        for fam, n in cases:
            C = dense_int(n, 20, seed=cluster).astype(np.float64) if fam == "ties" else gen.make_instance(fam, n, seed=11 + n)
            for sigma in (None, 0.0, 1e-2):
                u, v = mintrick_seeds(C, rng) if sigma is None else noisy_oracle_seeds(C, sigma)
                rco, xo, yo, co, tro = _oracle_seeded(C, u, v)
                Cd = torch.from_numpy(C.astype(np.float32)).cuda()
                x, y, rc, tr = ctx.solve_seeded(Cd, torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda(), want_trace=True)
                ctx.sync()
                assert int(rc[0]) == rco, (fam, n, sigma)
                if rco == 0:
                    assert np.array_equal(x[0].cpu().numpy(), xo) and np.array_equal(y[0].cpu().numpy(), yo), (fam, n, sigma)
                    t = tr[0].cpu().numpy()
                    assert (tro["aug_paths"], tro["collect_calls"], tro["relax_cols"], tro["arr_iters"]) == (t[7], t[8], t[9], t[6]), (fam, n, sigma)
            xc, yc, rcc = ctx.solve_cold(Cd)[:3]
            ctx.sync()
            xo, yo = oracle.port_lapjv_internal(C)
            assert int(rcc[0]) == 0 and np.array_equal(xc[0].cpu().numpy(), xo), (fam, n, "cold")
        # a batch: one cluster per instance
        n, B = 512, 6
        batch = gen.mixed_batch(n, B, first_seed=5)
        Cs = np.stack([c for _, c in batch])
        us, vs = zip(*[mintrick_seeds(c, rng) for c in Cs])
        x, y, rc = ctx.solve_seeded(torch.from_numpy(Cs.astype(np.float32)).cuda(), torch.from_numpy(np.stack(us)).cuda(), torch.from_numpy(np.stack(vs)).cuda())
        ctx.sync()
        for b in range(B):
            rco, xo, yo, co, _ = _oracle_seeded(Cs[b], us[b], vs[b])
            assert int(rc[b]) == rco and np.array_equal(x[b].cpu().numpy(), xo), batch[b][0]
    finally:
        ctx.set_option("solver_cluster", 0)


def test_large_instances_by_property(ctx):
    """Sizes the oracle does not finish quickly: permutation validity + strong duality of the final potentials."""
    import torch
    for fam, n in (("metric", 4096), ("uniform", 8192)):
        C = gen.make_instance(fam, n, seed=42)
        Cd = torch.from_numpy(C.astype(np.float32)).cuda()
        rng = np.random.default_rng(2)
        u, v = mintrick_seeds(C, rng)
        x, y, rc = ctx.solve_seeded(Cd, torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda())
        ctx.sync()
        assert int(rc[0]) == 0
        x = x[0].cpu().numpy().astype(np.int64); y = y[0].cpu().numpy().astype(np.int64)
        assert np.array_equal(np.sort(x), np.arange(n)) and np.array_equal(y[x], np.arange(n))
        cost = C[np.arange(n), x].sum()
        xc, yc, rcc, vfin = ctx.solve_cold(Cd, want_v=True)
        ctx.sync()
        xc = xc[0].cpu().numpy().astype(np.int64)
        vfin = vfin[0].cpu().numpy()
        ufin = np.min(C - vfin[None, :], axis=1)
        lower = ufin.sum() + vfin.sum()
        cold_cost = C[np.arange(n), xc].sum()
        assert abs(cold_cost - lower) <= 1e-9 * max(1.0, abs(lower)) * n    # cold solve is optimal (strong duality)
        assert abs(cost - cold_cost) <= 1e-9 * max(1.0, abs(cold_cost))     # and the seeded solve reaches the same optimum


# ---- dense half -------------------------------------------------------------------------------------
@pytest.mark.parametrize("fam", ("uniform", "sparse", "sparse1e6", "metric", "clustered", "odd", "tiny"))
def test_row_features_match_golden(golden, fam):
    from gnn.features import compute_row_features
    C = golden[f"{fam}/C"]
    f = compute_row_features(C)
    assert f.dtype == np.float32 and f.shape == (C.shape[0], 21)
    feature_close(f, golden[f"{fam}/feat"], rtol=1e-4)


def _pattern_rows(n, rng):
    """Rows built to defeat a sampled bracket: ties at / around the median, sorted and periodic rows, outliers."""
    P = rng.uniform(0, 1, (n, n))
    j = np.arange(n)
    P[0] = 3.25
    P[1] = np.where(j < n // 2, 0.0, 1.0)
    P[2] = j
    P[3] = j[::-1] * 0.5
    P[4] = j % 4
    P[5] = 1000.0; P[5, n // 3] = 0.0
    P[6] = rng.integers(0, 3, n)
    P[7] = np.where(j % 2 == 0, 5.0, rng.uniform(4.9, 5.1, n))
    P[8] = np.where(j % (n // 512 if n >= 1024 else 2) == 0, rng.uniform(0, 1, n), 7.0)      # period == the sample stride
    P[9] = rng.normal(50, 10, n)                                                           # isolated minimum: entropy ~ 0
    P[10] = np.exp(rng.normal(0, 4, n))                                                    # heavy tail: one histogram bin holds almost all
    P[11:40, :5] = 0.0                                                                     # ties at the row minimum
    return gen.snap_to_fp32_grid(P)


@pytest.mark.parametrize("n,opts", [(64, {}), (516, {}), (1030, {}), (2048, {}), (2048, {"feat_nbuf": 1, "feat_threads": 256}),
                                    (2048, {"feat_nsamp": 16}), (4096, {}), (8192, {"feat_threads": 512}), (16384, {})])
def test_streaming_row_features_patterns_and_options(ctx, n, opts):
    """features_smem.cuh on adversarial rows and non-default launch shapes (tiny sample => the bracket misses and
    the exact whole-row fall-back runs).  Order statistics exact, floating features within 1e-4."""
    import torch
    rng = np.random.default_rng(n)
    C = _pattern_rows(n, rng)
    rows = np.r_[0:40, n - 8:n] if n > 2048 else np.arange(n)
    ref = features_np.row_features(C)[rows] if n <= 2048 else None
    if ref is None:
        full = features_np.row_features(C[rows])          # row statistics of the chosen rows ...
        colmin = C.min(axis=0)
        full[:, 12] = (C[rows] == colmin[None, :]).mean(axis=1)                 # ... is_col_best needs all rows
        full[:, 13:] = features_np.positional_terms(n)[rows]
        ref = full
    try:
        for k, v in opts.items():
            ctx.set_option(k, v)
        feat, topv = ctx.row_features(torch.from_numpy(C.astype(np.float32)).cuda(), topk=16)
        ctx.sync()
    finally:
        for k in opts:
            ctx.set_option(k, 0)
    feature_close(feat[0].cpu().numpy()[rows], ref, rtol=1e-4)
    assert np.array_equal(topv[0].cpu().numpy()[rows], np.sort(C[rows].astype(np.float32), axis=1)[:, :16])


@pytest.mark.parametrize("fam", FAMILIES)
@pytest.mark.parametrize("n", [100, 512, 1030, 2048])
def test_row_features_match_numpy_definition(ctx, fam, n):
    import torch
    from gnn.features import compute_row_features
    C = gen.make_instance(fam, n, seed=42)
    ref = features_np.row_features(C)
    feature_close(compute_row_features(C), ref, rtol=1e-4)
    # device-pointer entry, binary32 and binary64 storage, and the top-k values
    for dt in (np.float32, np.float64):
        feat, topv = ctx.row_features(torch.from_numpy(C.astype(dt)).cuda(), topk=16)
        ctx.sync()
        feature_close(feat[0].cpu().numpy(), ref, rtol=1e-4)
        assert np.array_equal(topv[0].cpu().numpy(), np.sort(C.astype(np.float32), axis=1)[:, :16])


def test_small_model_matches_golden(ctx, golden):
    import torch
    import b200lap
    sd = state_dict_from_golden(golden)
    model = b200lap.Model(ctx, sd, topk=8)
    for fam in ("uniform", "sparse", "sparse1e6", "metric", "clustered", "odd", "tiny"):
        C = golden[f"{fam}/C"]
        Cd = torch.from_numpy(C.astype(np.float32)).cuda()
        feat, topv = ctx.row_features(Cd, topk=8)
        u, raw = ctx.onegnn_forward(model, feat, topv, want_raw=True)
        ctx.sync()
        scale = np.abs(golden[f"{fam}/small_raw"]).max()
        assert np.abs(raw[0].cpu().numpy() - golden[f"{fam}/small_raw"]).max() <= 1e-4 * scale, fam
        assert np.abs(u[0].cpu().numpy() - golden[f"{fam}/small_u"]).max() <= 1e-4 * scale, fam
        v = ctx.min_trick(Cd, torch.from_numpy(golden[f"{fam}/small_u"]).cuda())
        ctx.sync()
        assert np.array_equal(v[0].cpu().numpy(), golden[f"{fam}/small_v"]), fam     # exact given u


def test_named_model_matches_golden_and_module_mirror(ctx, golden):
    """hidden=192, layers=4, k=16 with torch.manual_seed(0) init, through the OneGNN module mirror."""
    import torch
    from gnn.one_gnn import OneGNN
    torch.manual_seed(0)
    model = OneGNN(21, hidden=192, layers=4, dropout=0.1, topk=16).eval()
    for fam in FAMILIES:
        C = golden[f"{fam}/C"]
        n = C.shape[0]
        row = torch.from_numpy(golden[f"{fam}/feat"]).float().unsqueeze(0)
        cost = torch.from_numpy(C).float().unsqueeze(0)
        mask = torch.ones((1, n), dtype=torch.bool)
        with torch.inference_mode():
            u = model(row, cost=cost, mask=mask)["u"].squeeze(0).cpu().numpy()
        scale = np.abs(golden[f"{fam}/named_raw"]).max()
        assert np.abs(u - golden[f"{fam}/named_u"]).max() <= 1e-4 * scale, fam
    # partial mask: masked rows are zeroed and contribute no message
    n = golden["uniform/C"].shape[0]
    mask = torch.ones((1, n), dtype=torch.bool)
    mask[0, ::3] = False
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    C = golden["uniform/C"]
    ref = onegnn_np.forward(sd, golden["uniform/feat"], cost=C.astype(np.float32), mask=mask[0].numpy())
    with torch.inference_mode():
        u = model(torch.from_numpy(golden["uniform/feat"]).unsqueeze(0), cost=torch.from_numpy(C).float().unsqueeze(0), mask=mask)["u"][0].cpu().numpy()
    assert np.abs(u - ref).max() <= 1e-4 * np.abs(golden["uniform/named_raw"]).max()
    assert (u[::3] == 0).all()


@pytest.mark.parametrize("fam", FAMILIES)
def test_pipeline_end_to_end(ctx, golden, fam):
    """predict (features -> OneGNN -> min-trick) then solve: u within tolerance of the CPU path, v exact given
    OUR u, and the assignment bit-exact against the oracle fed the same (C, u, v)."""
    import torch
    import b200lap
    torch.manual_seed(0)
    from gnn.one_gnn import OneGNN
    module = OneGNN(21, hidden=192, layers=4, dropout=0.1, topk=16).eval()
    sd = {k: v.detach().numpy() for k, v in module.state_dict().items()}
    pred = b200lap.GNNPredictor(module, device=0)
    n = 768
    C = gen.make_instance(fam, n, seed=42)
    u, v = pred.predict(C)
    u_ref, raw_ref = onegnn_np.forward(sd, features_np.row_features(C), cost=C.astype(np.float32), mask=np.ones(n, bool), return_raw=True)
    assert np.abs(u - u_ref.astype(np.float64)).max() <= 1e-4 * np.abs(raw_ref).max()
    assert np.array_equal(v, pipeline_np.min_trick(C, u.astype(np.float32)))
    import lap
    x, y, cost = lap.lapjv_seeded(C, u, v)
    xo, yo, co = oracle.port_lapjv_seeded(C, u, v)
    assert np.array_equal(x, xo) and np.array_equal(y, yo) and cost == co
    # device-resident chaining gives the same answer
    Cd = pred.to_device(C)
    xd, yd, rc, u64, v64 = ctx.pipeline(pred.model, Cd)
    ctx.sync()
    assert int(rc[0]) == 0 and np.array_equal(u64[0].cpu().numpy(), u) and np.array_equal(v64[0].cpu().numpy(), v)
    assert np.array_equal(xd[0].cpu().numpy(), xo) and np.array_equal(yd[0].cpu().numpy(), yo)


def test_min_trick_exact_at_full_size(ctx):
    """n = 16384 (1 GiB binary32): v equals a float64 recomputation from the same u, column block by column block."""
    import torch
    n = 16384
    g = torch.Generator(device="cuda").manual_seed(42)
    C = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
    u = (torch.randn(n, generator=g, device="cuda", dtype=torch.float32) * 0.01)
    v = ctx.min_trick(C, u)
    ctx.sync()
    ref = torch.empty(n, dtype=torch.float64, device="cuda")
    for j0 in range(0, n, 1024):
        ref[j0:j0 + 1024] = (C[:, j0:j0 + 1024].double() - u.double()[:, None]).min(dim=0).values
    assert torch.equal(v[0], ref)
    colmin, colarg = ctx.col_argmin(C)
    ctx.sync()
    m = C.min(dim=0)
    assert torch.equal(colmin[0], m.values)
    assert torch.equal(C[colarg[0].long(), torch.arange(n, device="cuda")], m.values)


def test_tensor_core_mlp_matches_ffma_path(ctx):
    """hidden=192 runs on tcgen05 (3xTF32, csrc/mlp_tc.cuh) by default; the FFMA tile kernel (csrc/mlp.cuh, ctx
    option mlp_impl=1) is the same math in plain binary32.  Both must agree to binary32 rounding level, with and
    without the cost branch, including a ragged last tile (n % 128 != 0) and a batch."""
    import torch
    import b200lap
    from gnn.one_gnn import OneGNN
    torch.manual_seed(0)
    module = OneGNN(21, hidden=192, layers=4, dropout=0.1, topk=16).eval()
    model = b200lap.Model(ctx, module.state_dict(), topk=16)
    for n, B in ((77, 1), (300, 3), (2048, 2)):
        Cs = np.stack([c for _, c in gen.mixed_batch(n, B, first_seed=9)]).astype(np.float32)
        Cd = torch.from_numpy(Cs).cuda()
        feat, topv = ctx.row_features(Cd, topk=16)
        for tv in (topv, None):
            ctx.set_option("mlp_impl", 1)
            u1, r1 = ctx.onegnn_forward(model, feat, tv, want_raw=True)
            ctx.set_option("mlp_impl", 0)
            u0, r0 = ctx.onegnn_forward(model, feat, tv, want_raw=True)
            ctx.sync()
            scale = float(r1.abs().max())
            assert float((r0 - r1).abs().max()) <= 2e-5 * scale, (n, B, tv is None)
            assert float((u0 - u1).abs().max()) <= 2e-5 * scale
