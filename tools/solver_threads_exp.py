"""Solver CTA size vs throughput on the bench batch: with 256-thread CTAs two instances share an SM (128 registers x 256
threads x 2, 2 x 85 KB of shared memory), so the barrier / replay phases of one overlap the relax steps of the other.
Prints per setting: one batch alone (ms) and the steady-state step with 8 batches in flight (ms)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from bench import named_state_dict, make_batch
ctx = b200lap.default_context(0)
model = b200lap.Model(ctx, named_state_dict(), topk=16)
Ch, fams = make_batch(0)
Cd = torch.from_numpy(Ch.astype(np.float32)).cuda()
stream = ctx.torch_stream()
u64, v64, _ = ctx.predict_duals(model, Cd)
ref = None
for T in (0, 256, 128, 1024):
    ctx.set_option("solver_threads", T)
    x, y, rc = ctx.solve_seeded(Cd, u64, v64); ctx.sync()
    if ref is None: ref = x.clone()
    same = bool(torch.equal(ref, x))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(3): ctx.solve_seeded(Cd, u64, v64)
    b.record(stream); ctx.sync()
    alone = a.elapsed_time(b) / 3
    res = {}
    for lanes in (4, 8):
        ctx.set_overlap(lanes)
        for _ in range(lanes): ctx.pipeline(model, Cd)
        ctx.sync()
        a.record(stream)
        K = 24
        for _ in range(K): ctx.pipeline(model, Cd)
        ctx.join(); b.record(stream); ctx.sync()
        res[lanes] = a.elapsed_time(b) / K
        ctx.set_overlap(False)
    print(f"solver_threads={T or 'auto(512)'}: solve alone {alone:.2f} ms, pipeline step with 4 lanes {res[4]:.2f} ms ({64/res[4]*1e3:.0f} inst/s), 8 lanes {res[8]:.2f} ms ({64/res[8]*1e3:.0f} inst/s), same={same}", flush=True)
ctx.set_option("solver_threads", 0)
