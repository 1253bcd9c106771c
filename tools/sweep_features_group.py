"""Times the group row-feature kernel (features_group.cuh) against the round-1 kernels (feat_impl=3: CTA streaming
kernel, feat_impl=4: warp-per-row) and checks they agree; prints the rows handed to the fall-back.
Usage: python tools/sweep_features_group.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from solvers import generators as gen

ctx = b200lap.default_context(0)
stream = ctx.torch_stream()
OPTS = ("feat_impl", "feat_group", "feat_ctas", "feat_stream")

def timed(fn, reps=10):
    fn(); fn(); ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    b.record(stream); ctx.sync()
    return a.elapsed_time(b) / reps

def setopts(**kw):
    for k in OPTS: ctx.set_option(k, kw.get(k, 0))

g = torch.Generator(device="cuda").manual_seed(42)
cases = [("n2048x64", torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(2048, 64, first_seed=42)]).astype(np.float32)).cuda(),
          [dict(feat_impl=4), dict(), dict(feat_stream=1), dict(feat_stream=2)]),
         ("n16384x1", torch.rand((16384, 16384), generator=g, device="cuda", dtype=torch.float32),
          [dict(feat_impl=3), dict(), dict(feat_group=8, feat_stream=2), dict(feat_group=4, feat_stream=1), dict(feat_group=2, feat_stream=1), dict(feat_group=8, feat_stream=1)]),
         ("n4096x16", torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(4096, 16, first_seed=42)]).astype(np.float32)).cuda(),
          [dict(feat_impl=3), dict(), dict(feat_group=2, feat_stream=2), dict(feat_group=1, feat_stream=1), dict(feat_group=2, feat_stream=1)]),
         ("n8192x4", torch.rand((4, 8192, 8192), generator=g, device="cuda", dtype=torch.float32),
          [dict(feat_impl=3), dict(), dict(feat_group=4, feat_stream=2), dict(feat_group=2, feat_stream=1), dict(feat_group=4, feat_stream=1)]),
         ("n512x256", torch.rand((256, 512, 512), generator=g, device="cuda", dtype=torch.float32), [dict(feat_impl=4), dict()]),
         ("n1024x128", torch.rand((128, 1024, 1024), generator=g, device="cuda", dtype=torch.float32), [dict(feat_impl=4), dict()])]
for name, C, grid in cases:
    f0 = t0 = None
    for o in grid:
        setopts(**o)
        try:
            ms = timed(lambda: ctx.row_features(C, topk=16))
            f, t = ctx.row_features(C, topk=16); ctx.sync()
            redo = ctx.feature_redo_rows() if o.get("feat_impl", 0) == 0 else -1
            if f0 is None: f0, t0 = f, t
            same = bool(torch.equal(t, t0))
            rel = float(((f - f0).abs() / (f0.abs() * 1e-4 + 1e-7)).max())
            print(f"{name} {str(o):40s} {ms:8.3f} ms {4.0*C.numel()/ms/1e6:8.1f} GB/s  redo_rows={redo} topk_equal={same} max_err/tol={rel:.3f}", flush=True)
        except Exception as e:
            print(name, o, "ERR", e, flush=True)
    del C
setopts()
for f in ("uniform", "sparse", "metric", "clustered"):
    C = torch.from_numpy(np.stack([gen.make_instance(f, 2048, seed=42 + k) for k in range(16)]).astype(np.float32)).cuda()
    for o in (dict(feat_impl=4), dict()):
        setopts(**o)
        ms = timed(lambda: ctx.row_features(C, topk=16))
        redo = ctx.feature_redo_rows() if not o else -1
        print(f"family {f:10s} n2048x16 {str(o):20s} {ms:8.3f} ms {4.0*C.numel()/ms/1e6:8.1f} GB/s redo_rows={redo}", flush=True)
setopts()
