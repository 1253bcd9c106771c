import os, sys
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from bench import named_state_dict
ctx = b200lap.default_context(0)
model = b200lap.Model(ctx, named_state_dict(), topk=16)
stream = ctx.torch_stream()
for n, B in ((8192, 4), (16384, 1)):
    g = torch.Generator(device="cuda").manual_seed(42)
    Cd = torch.rand((B, n, n), generator=g, device="cuda", dtype=torch.float32)
    u64, v64, _ = ctx.predict_duals(model, Cd)
    ref = None
    for T in (0, 512, 256):
        for nc in (8, 4):
            ctx.set_option("solver_threads", T); ctx.set_option("solver_cluster", nc)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = ctx.solve_seeded(Cd, u64, v64)
            e1.record(stream); ctx.sync()
            if ref is None: ref = out[0].clone()
            print(f"n={n} T={T} cluster={nc}: {e0.elapsed_time(e1):9.1f} ms same={bool(torch.equal(out[0], ref))}", flush=True)
