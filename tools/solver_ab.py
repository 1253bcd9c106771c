"""A/B of the solver's augmentation paths on the bench batch: register-resident d/v (solver_path.cuh) against the
shared-memory state path, a few CTA sizes each.  python tools/solver_ab.py [n B]"""
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
def timed(fn, reps=3):
    fn(); ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): out = fn()
    b.record(stream); ctx.sync()
    return a.elapsed_time(b) / reps, out
shapes = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(2048, 64), (512, 64), (4096, 16)]
for n, B in shapes:
    Cd = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(n, B, first_seed=42)]).astype(np.float32)).cuda()
    u64, v64, _ = ctx.predict_duals(model, Cd)
    ref = None
    for reg in (1, 0):
        for T in (0, 256, 512, 1024):
            if T and (T * 16 < n or T * 4 > n * 2): continue
            ctx.set_option("solver_regpath", reg)
            ctx.set_option("solver_threads", T)
            ms, out = timed(lambda: ctx.solve_seeded(Cd, u64, v64, want_trace=True))
            x = out[0]
            if ref is None: ref = x.clone()
            tr = out[3].cpu().numpy()
            print(f"n={n} B={B} regpath={reg} T={T or 'auto':>4}: {ms:8.2f} ms  same={bool(torch.equal(x, ref))} relax={int(tr[:,9].sum())} "
                  f"collects={int(tr[:,8].sum())} max_relax={int(tr[:,9].max())}", flush=True)
    ctx.set_option("solver_threads", 0)
    ctx.set_option("solver_regpath", 1)
