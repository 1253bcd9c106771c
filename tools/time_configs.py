"""Times the other BASELINE.json configs (parity-test shapes, not bench lines) on one B200 and prints JSON lines:
  1: single uniform n=512, host API latency        3: metric n=4096 x32 (fallback path)
  4: single uniform n=16384 (dense pass + solve)   5: n=8192 mixed, solver only, oracle duals + noise (reduced batch)
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from bench import named_state_dict
from solvers import generators as gen

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="1,3,4,5")
ap.add_argument("--batch5", type=int, default=8)
a = ap.parse_args()
ctx = b200lap.default_context(0)
model = b200lap.Model(ctx, named_state_dict(), topk=16)
stream = ctx.torch_stream()

def ev_time(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): out = fn()
    e1.record(stream); ctx.sync()
    return e0.elapsed_time(e1) / reps, out

def trace_summary(tr):
    tr = tr.cpu().numpy()
    return {"fallback": int(tr[:, 3].sum()), "paths": int(tr[:, 7].sum()), "relax_steps": int(tr[:, 9].sum()),
            "collects": int(tr[:, 8].sum()), "arr_iters": int(tr[:, 6].sum()), "proj_triggers": int(tr[:, 0].sum())}

cfgs = set(a.configs.split(","))
if "1" in cfgs:
    import lap
    from gnn.one_gnn import OneGNN
    C = gen.make_instance("uniform", 512, 42)
    pred = b200lap.GNNPredictor(named_state_dict(), device=0, ctx=ctx)
    def host():
        u, v = pred.predict(C)
        return lap.lapjv_seeded(C, u, v)
    host(); t0 = time.perf_counter()
    for _ in range(5): x, y, cost = host()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    Cd = pred.to_device(C)
    dms, out = ev_time(lambda: ctx.pipeline(model, Cd, want_trace=True), reps=10)
    print(json.dumps({"config": 1, "n": 512, "host_api_ms": round(ms, 3), "device_resident_ms": round(dms, 3), "cost": cost, "trace": trace_summary(out[5])}), flush=True)
if "3" in cfgs:
    n, B = 4096, 32
    Cs = np.stack([gen.make_instance("metric", n, 42 + k) for k in range(B)]).astype(np.float32)
    Cd = torch.from_numpy(Cs).cuda()
    ms, out = ev_time(lambda: ctx.pipeline(model, Cd, want_trace=True))
    pms, _ = ev_time(lambda: ctx.predict_duals(model, Cd))
    assert (out[2] == 0).all()
    print(json.dumps({"config": 3, "n": n, "batch": B, "pipeline_ms": round(ms, 3), "predict_ms": round(pms, 3), "inst_per_s": round(B / ms * 1e3, 1), "trace": trace_summary(out[5])}), flush=True)
    del Cd
if "4" in cfgs:
    n = 16384
    g = torch.Generator(device="cuda").manual_seed(42)
    Cd = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
    pms, duals = ev_time(lambda: ctx.predict_duals(model, Cd), reps=5)
    u64, v64, _ = duals
    t0 = time.perf_counter()
    x, y, rc, tr = ctx.solve_seeded(Cd, u64, v64, want_trace=True); ctx.sync()
    sms = (time.perf_counter() - t0) * 1e3
    assert int(rc[0]) == 0
    xs = x[0].long()
    assert bool((torch.sort(xs).values == torch.arange(n, device="cuda")).all())
    print(json.dumps({"config": 4, "n": n, "predict_ms": round(pms, 3), "dense_pass_frac_of_one_read_roofline": round((4.0 * n * n / (pms * 1e-3)) / 6546.2e9, 4),
                      "solve_ms": round(sms, 1), "inst_per_s": round(1e3 / (pms + sms), 4), "trace": trace_summary(tr)}), flush=True)
    del Cd
if "5" in cfgs:
    n, B = 8192, a.batch5
    Cs = np.stack([c for _, c in gen.mixed_batch(n, B, first_seed=42)]).astype(np.float32)
    Cd = torch.from_numpy(Cs).cuda()
    t0 = time.perf_counter()
    xc, yc, rcc, vfin = ctx.solve_cold(Cd, want_v=True); ctx.sync()
    cold_ms = (time.perf_counter() - t0) * 1e3
    u_opt = (Cd.double() - vfin[:, None, :]).min(dim=2).values
    for sigma in (0.0, 1e-3):
        gn = torch.Generator(device="cuda").manual_seed(42)
        u = u_opt + sigma * torch.randn(u_opt.shape, generator=gn, device="cuda", dtype=torch.float64)
        v = vfin + sigma * torch.randn(u_opt.shape, generator=gn, device="cuda", dtype=torch.float64)
        t0 = time.perf_counter()
        x, y, rc, tr = ctx.solve_seeded(Cd, u, v, want_trace=True); ctx.sync()
        ms = (time.perf_counter() - t0) * 1e3
        assert (rc == 0).all()
        same_cost = bool(torch.allclose(torch.gather(Cd, 2, x.long().unsqueeze(-1)).sum(dim=(1, 2)).double(), torch.gather(Cd, 2, xc.long().unsqueeze(-1)).sum(dim=(1, 2)).double(), rtol=1e-6))
        print(json.dumps({"config": 5, "n": n, "batch": B, "sigma": sigma, "solve_ms": round(ms, 1), "inst_per_s": round(B / ms * 1e3, 3), "cold_ms": round(cold_ms, 1),
                          "optimal_cost": same_cost, "trace": trace_summary(tr)}), flush=True)
