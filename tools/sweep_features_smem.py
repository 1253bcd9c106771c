"""Times the streaming row-feature kernel (features_smem.cuh) over its launch options and against the
register-resident kernel (feat_impl=1); checks the two agree.  Usage: python tools/sweep_features_smem.py [quick]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from solvers import generators as gen

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
ctx = b200lap.default_context(0)
stream = ctx.torch_stream()
OPTS = ("feat_impl", "feat_threads", "feat_nbuf", "feat_nsamp", "feat_ctas")

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
Cb = torch.rand((16384, 16384), generator=g, device="cuda", dtype=torch.float32)
Cs = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(2048, 64, first_seed=42)]).astype(np.float32)).cuda()
C4 = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(4096, 16, first_seed=42)]).astype(np.float32)).cuda()
C8 = torch.rand((4, 8192, 8192), generator=g, device="cuda", dtype=torch.float32)
fam = {}
for i, f in enumerate(("uniform", "sparse", "metric", "clustered")):
    fam[f] = torch.from_numpy(np.stack([gen.make_instance(f, 2048, seed=42 + k) for k in range(16)]).astype(np.float32)).cuda()

cases = [("n16384x1", Cb), ("n2048x64", Cs), ("n4096x16", C4), ("n8192x4", C8)]
for name, C in cases:
    setopts(feat_impl=1)
    ms0 = timed(lambda: ctx.row_features(C, topk=16), reps=3)
    f0, t0 = ctx.row_features(C, topk=16); ctx.sync()
    print(f"{name} register-resident kernel      {ms0:8.3f} ms {4.0*C.numel()/ms0/1e6:8.1f} GB/s", flush=True)
    n = C.shape[-1]
    grid = [dict()]
    if not quick:
        for T in (128, 256, 512):
            for nbuf in (1, 2):
                grid.append(dict(feat_threads=T, feat_nbuf=nbuf))
        for s in (256, 512, 1024, 2048, 4096):
            grid.append(dict(feat_nsamp=s))
        for c in (1, 2, 3, 4, 6):
            grid.append(dict(feat_ctas=c))
    for o in grid:
        setopts(**o)
        try:
            ms = timed(lambda: ctx.row_features(C, topk=16))
            f, t = ctx.row_features(C, topk=16); ctx.sync()
            same = bool(torch.equal(t, t0))
            rel = float(((f - f0).abs() / (f0.abs() * 1e-4 + 1e-7)).max())
            print(f"{name} streaming {str(o):48s} {ms:8.3f} ms {4.0*C.numel()/ms/1e6:8.1f} GB/s  topk_equal={same} max_err/tol={rel:.3f}", flush=True)
        except Exception as e:
            print(name, o, "ERR", e, flush=True)
setopts()
for f, C in fam.items():
    ms = timed(lambda: ctx.row_features(C, topk=16))
    print(f"family {f:10s} n2048x16 streaming {ms:8.3f} ms {4.0*C.numel()/ms/1e6:8.1f} GB/s", flush=True)
