"""A/B of solver options on fixed seeds: python tools/ab_solver.py"""
import os, sys
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
def timed(fn, reps=4):
    fn(); ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): out = fn()
    b.record(stream); ctx.sync()
    return a.elapsed_time(b) / reps, out
for n, B in ((2048, 64), (4096, 16)):
    Cd = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(n, B, first_seed=42)]).astype(np.float32)).cuda()
    for impl in (1, 0):
        ctx.set_option("feat_impl", impl)
        u64, v64, _ = ctx.predict_duals(model, Cd)
        ref = None
        for _rep in (0, 1):

            ms, out = timed(lambda: ctx.solve_seeded(Cd, u64, v64, want_trace=True))
            x = out[0]
            if ref is None: ref = x.clone()
            tr = out[3].cpu().numpy()
            print(f"n={n} B={B} feat_impl={impl}: {ms:8.2f} ms same={bool(torch.equal(x, ref))} relax={int(tr[:,9].sum())} max_relax={int(tr[:,9].max())} collects={int(tr[:,8].sum())} paths={int(tr[:,7].sum())}", flush=True)
    ctx.set_option("feat_impl", 0)
