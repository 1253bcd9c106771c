"""One row-feature call per shape, for ncu: python tools/profile_features_group.py <n> <batch> [feat_group] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from solvers import generators as gen
n, batch = int(sys.argv[1]), int(sys.argv[2])
ctx = b200lap.default_context(0)
if len(sys.argv) > 3: ctx.set_option("feat_group", int(sys.argv[3]))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
if n <= 4096:
    C = torch.from_numpy(np.stack([c for _, c in gen.mixed_batch(n, batch, first_seed=42)]).astype(np.float32)).cuda()
else:
    C = torch.rand((batch, n, n), generator=torch.Generator(device="cuda").manual_seed(42), device="cuda", dtype=torch.float32)
for _ in range(reps):
    f, t = ctx.row_features(C, topk=16)
ctx.sync()
print("redo rows", ctx.feature_redo_rows(), "of", n * batch)
