"""Times the seeded solve of large instances for every cluster size (ctx option solver_cluster; 1 = single CTA)."""
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
sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]
for n in sizes:
    B = {4096: 8, 8192: 4, 16384: 1}.get(n, 1)
    g = torch.Generator(device="cuda").manual_seed(42)
    Cd = torch.rand((B, n, n), generator=g, device="cuda", dtype=torch.float32)
    u64, v64, _ = ctx.predict_duals(model, Cd)
    ref = None
    for nc in ((8, 4, 2, 1) if n < 16384 else (8, 4, 2, 1)):
        ctx.set_option("solver_cluster", nc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = ctx.solve_seeded(Cd, u64, v64, want_trace=True)
        e1.record(stream); ctx.sync()
        ms = e0.elapsed_time(e1)
        x = out[0]
        if ref is None: ref = x.clone()
        tr = out[3].cpu().numpy()
        print(f"n={n} B={B} cluster={nc}: {ms:10.1f} ms same={bool(torch.equal(x, ref))} rc={out[2].cpu().tolist()} relax={int(tr[:,9].sum())} collects={int(tr[:,8].sum())} paths={int(tr[:,7].sum())}", flush=True)
    ctx.set_option("solver_cluster", 0)
