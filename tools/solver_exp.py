"""Solver experiments on the bench batch: scans per batched relax step, pipelined hit replay, binary64-stored matrix.
python tools/solver_exp.py   (B200LAP_LIB_VARIANT=k8 for the 8-scan build)"""
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
n, B = 2048, 64
Cd = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(n, B, first_seed=42)]).astype(np.float32)).cuda()
u64, v64, _ = ctx.predict_duals(model, Cd)
ref = None
var = os.environ.get("B200LAP_LIB_VARIANT", "default")
for pipe in (1, 0):
    for kcap in (1, 2, 3, 4, 6, 8):
        if kcap > 4 and var != "k8": continue
        if kcap > 2 and var == "k2": continue
        ctx.set_option("solver_pipe", pipe); ctx.set_option("solver_kcap", kcap)
        ms, out = timed(lambda: ctx.solve_seeded(Cd, u64, v64, want_trace=True))
        if ref is None: ref = out[0].clone()
        print(f"lib={var} pipe={pipe} kcap={kcap}: {ms:8.2f} ms same={bool(torch.equal(out[0], ref))}", flush=True)
ctx.set_option("solver_pipe", 1); ctx.set_option("solver_kcap", 0)
C64 = Cd.double()
ms, out = timed(lambda: ctx.solve_seeded(C64, u64, v64, want_trace=True))
print(f"lib={var} binary64-stored matrix: {ms:8.2f} ms same={bool(torch.equal(out[0], ref))}", flush=True)
