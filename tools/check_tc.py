"""Compares the tcgen05 OneGNN path with the FFMA path (ctx option mlp_impl=1) on the same inputs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from bench import named_state_dict
from solvers import generators as gen

ctx = b200lap.default_context(0)
model = b200lap.Model(ctx, named_state_dict(), topk=16)
stream = ctx.torch_stream()
for n, B in ((300, 1), (2048, 2), (16384, 1)):
    if n == 16384:
        g = torch.Generator(device="cuda").manual_seed(1)
        C = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
    else:
        C = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(n, B, first_seed=5)]).astype(np.float32)).cuda()
    feat, topv = ctx.row_features(C, topk=16)
    for has_cost in (True, False):
        tv = topv if has_cost else None
        ctx.set_option("mlp_impl", 1)
        u_ref, raw_ref = ctx.onegnn_forward(model, feat, tv, want_raw=True); ctx.sync()
        ctx.set_option("mlp_impl", 0)
        u_tc, raw_tc = ctx.onegnn_forward(model, feat, tv, want_raw=True); ctx.sync()
        scale = float(raw_ref.abs().max())
        err = float((raw_tc - raw_ref).abs().max())
        print(f"n={n} B={B} cost={has_cost}: max|raw_tc-raw_ffma|={err:.3e} scale={scale:.3e} rel={err/scale:.3e} nan={bool(torch.isnan(raw_tc).any())}", flush=True)
    def timed(fn, reps=10):
        fn(); ctx.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps): fn()
        b.record(stream); ctx.sync()
        return a.elapsed_time(b) / reps
    ctx.set_option("mlp_impl", 1); t1 = timed(lambda: ctx.onegnn_forward(model, feat, topv))
    ctx.set_option("mlp_impl", 0); t0 = timed(lambda: ctx.onegnn_forward(model, feat, topv))
    print(f"n={n} B={B}: ffma {t1:.3f} ms, tcgen05 {t0:.3f} ms", flush=True)
