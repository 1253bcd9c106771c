"""One row-feature launch per shape, for ncu: python tools/profile_features.py [n batch]..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import torch
import b200lap

ctx = b200lap.default_context(0)
args = [int(a) for a in sys.argv[1:]] or [16384, 1, 2048, 64]
g = torch.Generator(device="cuda").manual_seed(42)
for n, b in zip(args[::2], args[1::2]):
    C = torch.rand((b, n, n), generator=g, device="cuda", dtype=torch.float32)
    for _ in range(2):
        f, t = ctx.row_features(C, topk=16)
    ctx.sync()
    print(n, b, float(f.sum()))
