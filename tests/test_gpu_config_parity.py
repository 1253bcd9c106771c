"""GPU parity at the sizes of BASELINE.json's configs, against the UNMODIFIED reference solver binary
(oracle/_ref/libreflap.so, compiled from /root/reference/LAP/_lapjv_cpp/{lapjv_seeded,lapjv}.cpp by oracle/Makefile;
it travels to the GPU box prebuilt).  Both implementations get the same (C, u, v); x, y and cost must be equal --
tie-breaking included -- and the ten phase counters must equal those of the instrumented C port (oracle/jv_port.c,
itself pinned to the reference binary).  The CPU side runs in a thread pool (ctypes releases the GIL).

  config 1   single uniform n = 512, full pipeline                 (also smoke())
  config 2   the whole 64-instance mixed-family n = 2048 bench batch, full pipeline
  config 3   32 x metric n = 4096, full pipeline (top-k refinement + full-JV fallback)
  config 4   single uniform n = 16384, dense pass + seeded solve in the default cluster mode
  config 5   n = 8192, one instance per family x sigma in {0, 1e-3, 1e-2}: optimal duals + noise, solver only,
             default cluster mode (seed law /root/reference/solvers/dual_computation.py:77-115)
plus the front-end sweep against the port's front end and the solvers/advanced_dual.py drop-ins against their NumPy
statements.  Nothing here reads /root/reference."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle
from solvers import generators as gen
from _fixtures import mintrick_seeds, noisy_oracle_seeds
import _capi

pytestmark = pytest.mark.gpu

COUNTERS = ("proj_triggers", "tight_edges", "greedy_matched", "took_fallback", "micro_bumps", "free_after_cr",
            "arr_iters", "aug_paths", "collect_calls", "relax_cols")


@pytest.fixture(scope="module")
def ctx():
    import b200lap
    return b200lap.default_context(0)


@pytest.fixture(scope="module")
def model(ctx):
    import b200lap
    from bench import named_state_dict
    return b200lap.Model(ctx, named_state_dict(), topk=16)


@pytest.fixture(scope="module")
def pool():
    with ThreadPoolExecutor(max_workers=max(2, min(32, os.cpu_count() or 2))) as ex:
        yield ex


def _need_ref():
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libreflap.so did not travel")


def _ref_job(C, u, v):
    return oracle.ref_lapjv_seeded(C, u, v)


def _port_job(C, u, v):
    tr = oracle.Trace()
    x, y, cost = oracle.port_lapjv_seeded(C, u, v, trace=tr)
    return x, y, cost, tr.as_dict()


def _check_batch(pool, Cs, us, vs, xs, ys, trs, tag, with_port=True):
    """(x, y, cost) against the reference binary and the ten counters against the port, instance by instance."""
    refs = [pool.submit(_ref_job, Cs[b], us[b], vs[b]) for b in range(len(Cs))]
    ports = [pool.submit(_port_job, Cs[b], us[b], vs[b]) for b in range(len(Cs))] if with_port else None
    for b in range(len(Cs)):
        xr, yr, cr = refs[b].result()
        assert np.array_equal(xs[b], xr) and np.array_equal(ys[b], yr), (tag, b, "assignment differs from the reference binary")
        n = Cs[b].shape[0]
        assert float(np.sum(Cs[b][np.arange(n), xs[b]])) == cr, (tag, b, "cost")
        if with_port:
            xp, yp, cp, trp = ports[b].result()
            assert np.array_equal(xp, xr) and cp == cr, (tag, b, "port vs reference")
            got = {k: int(trs[b][i]) for i, k in enumerate(_capi.TRACE)}
            bad = [(k, trp[k], got[k]) for k in COUNTERS if trp[k] != got[k]]
            assert not bad, (tag, b, bad)


def test_config2_whole_bench_batch_bit_exact(ctx, model, pool):
    import torch
    _need_ref()
    from bench import make_batch
    Ch, fams = make_batch(0)
    Cd = torch.from_numpy(Ch.astype(np.float32)).cuda()
    x, y, rc, u64, v64, tr = ctx.pipeline(model, Cd, want_trace=True)
    ctx.sync()
    assert (rc == 0).all()
    _check_batch(pool, Ch, u64.cpu().numpy(), v64.cpu().numpy(), x.cpu().numpy().astype(np.int64), y.cpu().numpy().astype(np.int64),
                 tr.cpu().numpy(), "config2")


def test_config1_single_uniform_512(ctx, model, pool):
    import torch
    _need_ref()
    C = gen.make_instance("uniform", 512, seed=42)
    x, y, rc, u64, v64, tr = ctx.pipeline(model, torch.from_numpy(C.astype(np.float32)).cuda(), want_trace=True)
    ctx.sync()
    assert int(rc[0]) == 0
    _check_batch(pool, [C], u64.cpu().numpy(), v64.cpu().numpy(), x.cpu().numpy().astype(np.int64), y.cpu().numpy().astype(np.int64),
                 tr.cpu().numpy(), "config1")


def test_config3_metric_4096_batch_bit_exact(ctx, model, pool):
    import torch
    _need_ref()
    B, n = 32, 4096
    Cs = [gen.make_instance("metric", n, seed=42 + k) for k in range(B)]
    Cd = torch.from_numpy(np.stack(Cs).astype(np.float32)).cuda()
    x, y, rc, u64, v64, tr = ctx.pipeline(model, Cd, want_trace=True)
    ctx.sync()
    assert (rc == 0).all()
    trn = tr.cpu().numpy()
    assert (trn[:, 3] == 1).all(), "the metric family takes the 1.2 n fallback"
    _check_batch(pool, Cs, u64.cpu().numpy(), v64.cpu().numpy(), x.cpu().numpy().astype(np.int64), y.cpu().numpy().astype(np.int64), trn, "config3")


def test_config5_n8192_noisy_oracle_duals_default_cluster_mode(ctx, pool):
    import torch
    _need_ref()
    n = 8192
    fams = ("uniform", "sparse", "metric", "clustered")
    Cs = [gen.make_instance(f, n, seed=42 + k) for k, f in enumerate(fams)]
    Cd = torch.from_numpy(np.stack(Cs).astype(np.float32)).cuda()
    # optimal duals from a cold JV run: v final, u_i = C[i, x_i] - v[x_i]
    xc, yc, rcc, vfin = ctx.solve_cold(Cd, want_v=True)
    ctx.sync()
    assert (rcc == 0).all()
    xc = xc.cpu().numpy().astype(np.int64); vfin = vfin.cpu().numpy()
    cases_C, cases_u, cases_v = [], [], []
    for sigma in (0.0, 1e-3, 1e-2):
        for b in range(len(fams)):
            u = Cs[b][np.arange(n), xc[b]] - vfin[b][xc[b]]
            v = vfin[b].copy()
            if sigma > 0:
                rs = np.random.RandomState(42)
                u = u + rs.normal(0, sigma, n)
                v = v + rs.normal(0, sigma, n)
            cases_C.append(b); cases_u.append(u); cases_v.append(v)
    idx = torch.tensor(cases_C, device="cuda")
    xs, ys, trs = [], [], []
    for q0 in range(0, len(cases_C), 4):                       # four instances per launch (2 GiB of matrices on the device)
        sel = slice(q0, q0 + 4)
        x, y, rc, tr = ctx.solve_seeded(Cd[idx[sel]], torch.from_numpy(np.stack(cases_u[sel])).cuda(),
                                        torch.from_numpy(np.stack(cases_v[sel])).cuda(), want_trace=True)
        ctx.sync()
        assert (rc == 0).all()
        xs += list(x.cpu().numpy().astype(np.int64)); ys += list(y.cpu().numpy().astype(np.int64)); trs += list(tr.cpu().numpy())
    _check_batch(pool, [Cs[b] for b in cases_C], cases_u, cases_v, xs, ys, trs, "config5")


def test_config4_uniform_16384_default_cluster_mode(ctx, model, pool):
    import torch
    _need_ref()
    n = 16384
    C = gen.make_instance("uniform", n, seed=42)
    Cd = torch.from_numpy(C.astype(np.float32)).cuda()
    x, y, rc, u64, v64, tr = ctx.pipeline(model, Cd, want_trace=True)
    ctx.sync()
    assert int(rc[0]) == 0
    _check_batch(pool, [C], u64.cpu().numpy(), v64.cpu().numpy(), x.cpu().numpy().astype(np.int64), y.cpu().numpy().astype(np.int64),
                 tr.cpu().numpy(), "config4", with_port=False)      # one CPU solve of this size is enough (about a minute)


# ---- the front-end sweep against the port's front end (lapjv_seeded.cpp:38-113) ---------------------------
@pytest.mark.parametrize("fam", ("uniform", "sparse", "sparse1e6", "metric", "clustered"))
def test_front_end_sweep_matches_port(ctx, fam):
    import torch
    rng = np.random.default_rng(5)
    for n in (96, 1000, 2048):
        C = gen.make_instance(fam, n, seed=3 + n)
        Cd = torch.from_numpy(C.astype(np.float32)).cuda()
        # seeds of the shape the dense half produces: no projection trigger, the sweep IS the front end
        u, v = mintrick_seeds(C, rng)
        ut, tc, viol, infeas, total = ctx.front_end(Cd, torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda())
        rc, uo, vo, xo, yo, tro = oracle.port_front_end(C, u, v)
        assert rc == 0 and tro["proj_triggers"] == 0
        assert not bool(viol[0]) and not bool(infeas[0])
        assert np.array_equal(ut[0].cpu().numpy(), uo), (fam, n, "u_tight")
        assert int(total[0]) == tro["tight_edges"] == int(tc[0].sum()), (fam, n, "tight count")
        # noisy optimal duals: the sweep must report the violation (the solver then redoes the front end in order)
        u2, v2 = noisy_oracle_seeds(C, 1e-2)
        _, _, viol2, _, _ = ctx.front_end(Cd, torch.from_numpy(u2).cuda(), torch.from_numpy(v2).cuda())
        rc2, _, _, _, _, tro2 = oracle.port_front_end(C, u2, v2)
        assert bool(viol2[0]) == (tro2["proj_triggers"] > 0), (fam, n, "violation flag")
        # ... and the solver's counters for that case equal the port's, projection triggers included
        x, y, rcs, tr = ctx.solve_seeded(Cd, torch.from_numpy(u2).cuda(), torch.from_numpy(v2).cuda(), want_trace=True)
        ctx.sync()
        trp = oracle.Trace()
        xo, yo, _ = oracle.port_lapjv_seeded(C, u2, v2, trace=trp)
        assert np.array_equal(x[0].cpu().numpy(), xo)
        got = {k: int(tr[0][i]) for i, k in enumerate(_capi.TRACE)}
        bad = [(k, trp.as_dict()[k], got[k]) for k in COUNTERS if trp.as_dict()[k] != got[k]]
        assert not bad, (fam, n, bad)


# ---- solvers/advanced_dual.py drop-ins against their NumPy statements (reference :14-63) ----------------------
def _np_project_feasible(C, u, v, max_rounds=50, tol=1e-12):
    u = u.copy(); v = v.copy()
    for _ in range(max(1, int(max_rounds))):
        u = np.minimum(u, (C - v[None, :]).min(axis=1))
        v = np.minimum(v, (C - u[:, None]).min(axis=0))
        if (C - u[:, None] - v[None, :]).min() >= -tol:
            break
    return u, v


@pytest.mark.parametrize("fam", ("uniform", "sparse1e6", "clustered"))
def test_advanced_dual_dropins(fam):
    from solvers import advanced_dual as ad
    rng = np.random.default_rng(12)
    for n in (64, 1000):
        for f64 in (False, True):
            C = gen.make_instance(fam, n, seed=9) if not f64 else rng.uniform(0, 1, (n, n))
            u, v = noisy_oracle_seeds(C, 5e-2)
            up, vp = ad.project_feasible(C, u, v)
            ur, vr = _np_project_feasible(C, u, v)
            assert np.array_equal(up, ur) and np.array_equal(vp, vr), (fam, n, f64)
            up1, vp1 = ad.project_feasible(C, u, v, max_rounds=1)
            ur1, vr1 = _np_project_feasible(C, u, v, max_rounds=1)
            assert np.array_equal(up1, ur1) and np.array_equal(vp1, vr1)
            assert ad.check_dual_feasible(C, up, vp) is True
            with pytest.raises(AssertionError):
                ad.check_dual_feasible(C, u + 1.0, v)
            red = C - up[:, None] - vp[None, :]
            assert ad.min_reduced_cost(C, up, vp) == red.min()
            R = ad.reduce_costs(C, u, v)                       # infeasible seeds: shifted by the (negative) minimum
            raw = C - u[:, None] - v[None, :]
            assert raw.min() < 0 and np.array_equal(R, raw - raw.min())
            assert np.array_equal(ad.reduce_costs(C, u, v, shift_nonneg=False), raw)
            assert np.array_equal(ad.reduce_costs(C, up, vp), red if red.min() >= 0 else red - red.min())


def test_batches_in_flight_on_every_lane_agree_with_one_at_a_time(ctx, model):
    """Context.set_overlap(L): whole-pipeline calls rotate over L lanes.  Every call must return what the
    same call returns alone -- also when the caller drops results while later lanes are still queued (the returned
    tensors carry the lane's stream, so torch's allocator cannot recycle them under a running lane)."""
    import torch
    fams = ("uniform", "sparse", "metric", "clustered")
    n, B = 512, 16
    C = torch.from_numpy(np.stack([gen.make_instance(fams[k % 4], n, seed=900 + k) for k in range(B)]).astype(np.float32)).cuda()
    ctx.set_overlap(False)
    x0, y0, rc0, u0, v0 = ctx.pipeline(model, C)
    ctx.sync()
    assert (rc0.cpu().numpy() == 0).all()
    x0, y0, u0, v0 = x0.clone(), y0.clone(), u0.clone(), v0.clone()
    for lanes in (2, 8):
        ctx.set_overlap(lanes)
        try:
            kept = []
            for step in range(3 * lanes):
                out = ctx.pipeline(model, C)
                kept.append(out if step % 3 == 0 else (out[0], out[1]))      # most calls drop rc / u / v at once
                del out
            ctx.sync()
            for k, o in enumerate(kept):
                assert torch.equal(o[0], x0) and torch.equal(o[1], y0), (lanes, k)
                if len(o) > 2:
                    assert (o[2].cpu().numpy() == 0).all() and torch.equal(o[3], u0) and torch.equal(o[4], v0), (lanes, k)
        finally:
            ctx.set_overlap(False)
