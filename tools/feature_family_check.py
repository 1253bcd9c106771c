import os, sys
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from solvers import generators as gen
ctx = b200lap.default_context(0)
stream = ctx.torch_stream()
def timed(fn, reps=5):
    fn(); fn(); ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    b.record(stream); ctx.sync()
    return a.elapsed_time(b) / reps
for fam in ("metric", "uniform", "clustered", "sparse"):
    for n, B in ((4096, 8), (8192, 2)):
        C = torch.from_numpy(np.stack([gen.make_instance(fam, n, seed=42 + k) for k in range(B)]).astype(np.float32)).cuda()
        for o in (dict(), dict(feat_stream=2, feat_group=(2 if n == 4096 else 4)), dict(feat_stream=1, feat_group=2)):
            for k in ("feat_stream", "feat_group"): ctx.set_option(k, o.get(k, 0))
            ms = timed(lambda: ctx.row_features(C, topk=16))
            print(fam, n, B, o, f"{ms:.3f} ms redo", ctx.feature_redo_rows(), flush=True)
        del C
