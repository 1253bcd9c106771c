"""Times k_row_features for every entries-per-thread variant (ctx option feat_ept) and the front end."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from solvers import generators as gen

ctx = b200lap.default_context(0)
stream = ctx.torch_stream()

def timed(fn, reps=5):
    fn(); ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    b.record(stream); ctx.sync()
    return a.elapsed_time(b) / reps

g = torch.Generator(device="cuda").manual_seed(42)
Cb = torch.rand((16384, 16384), generator=g, device="cuda", dtype=torch.float32)
Cs = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(2048, 64, first_seed=42)]).astype(np.float32)).cuda()
C4 = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(4096, 8, first_seed=42)]).astype(np.float32)).cuda()
for name, C in (("n16384x1", Cb), ("n2048x64", Cs), ("n4096x8", C4)):
    ref = None
    for ept in (4, 8, 16, 32, 64):
        ctx.set_option("feat_ept", ept)
        try:
            ms = timed(lambda: ctx.row_features(C, topk=16))
            f, t = ctx.row_features(C, topk=16); ctx.sync()
            if ref is None: ref = (f.clone(), t.clone())
            same = bool(torch.equal(t, ref[1])) and float((f - ref[0]).abs().max()) < 1e-3
            print(f"{name} ept={ept:2d} row_features {ms:8.3f} ms  {4.0*C.numel()/ms/1e6:8.1f} GB/s  consistent={same}", flush=True)
        except Exception as e:
            print(name, ept, "ERR", e, flush=True)
    ctx.set_option("feat_ept", 0)
    u = torch.zeros(C.shape[:-1] if C.dim()==3 else (1, C.shape[0]), dtype=torch.float64, device="cuda")
    v = C.min(dim=-2).values.double().reshape(u.shape)
    ms = timed(lambda: ctx.front_end(C, u, v))
    print(f"{name} front_end {ms:8.3f} ms {4.0*C.numel()/ms/1e6:8.1f} GB/s", flush=True)
