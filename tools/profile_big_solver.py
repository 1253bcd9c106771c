"""One cold solve of a single uniform n=8192 instance (the ARR-dominated fallback path) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, time
import b200lap
from solvers import generators as gen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = b200lap.default_context(0)
C = torch.from_numpy(gen.make_instance("uniform", n, 42).astype(np.float32)).cuda()
t0 = time.perf_counter()
x, y, rc, tr = ctx.solve_cold(C, want_trace=True); ctx.sync()
print("cold solve n=%d: %.1f ms" % (n, (time.perf_counter() - t0) * 1e3), tr.cpu().numpy()[0][:10])
