import os, sys
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from solvers import generators as gen
ctx = b200lap.default_context(0); stream = ctx.torch_stream()
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
Cs = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(2048, 64, first_seed=42)]).astype(np.float32)).cuda()
C4 = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(4096, 16, first_seed=42)]).astype(np.float32)).cuda()
Cb = torch.rand((16384, 16384), generator=g, device="cuda", dtype=torch.float32)
for name, C in (("n2048x64", Cs), ("n4096x16", C4), ("n16384", Cb)):
    setopts(); f0, t0 = ctx.row_features(C, topk=16); ctx.sync()
    for T in (32, 64, 128, 256):
        for nbuf in (1, 2):
            for s in (0, 256):
                o = dict(feat_threads=T, feat_nbuf=nbuf, feat_nsamp=s)
                setopts(**o)
                try:
                    ms = timed(lambda: ctx.row_features(C, topk=16))
                    f, t = ctx.row_features(C, topk=16); ctx.sync()
                    rel = float(((f - f0).abs() / (f0.abs() * 1e-4 + 1e-7)).max())
                    print(f"{name} {str(o):60s} {ms:8.3f} ms {4.0*C.numel()/ms/1e6:8.1f} GB/s eq={bool(torch.equal(t,t0))} err/tol={rel:.3f}", flush=True)
                except Exception as e:
                    print(name, o, "ERR", str(e)[:100], flush=True)
