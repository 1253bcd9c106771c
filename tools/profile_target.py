"""Small driver for ncu captures (run on the GPU box): one warm pass then `--reps` passes of the dense
kernels at n=16384 and of the n=2048 pipeline at a small batch, so `ncu -k regex:... -s ... -c ...` sees
steady-state launches.  Prints nothing that is a bench number."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import b200lap  # noqa: E402
from bench import named_state_dict  # noqa: E402
from solvers import generators as gen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-big", type=int, default=16384)
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--what", default="dense,pipeline")
    a = ap.parse_args()
    ctx = b200lap.default_context(0)
    model = b200lap.Model(ctx, named_state_dict(), topk=16)
    if "dense" in a.what:
        g = torch.Generator(device="cuda").manual_seed(42)
        Cb = torch.rand((a.n_big, a.n_big), generator=g, device="cuda", dtype=torch.float32)
        for _ in range(1 + a.reps):
            u64, v64, u32 = ctx.predict_duals(model, Cb)
            ctx.front_end(Cb, u64, v64)
        ctx.sync()
        del Cb
    if "pipeline" in a.what:
        Cs = np.stack([c for _, c in gen.mixed_batch(a.n, a.batch, first_seed=42)]).astype(np.float32)
        Cd = torch.from_numpy(Cs).cuda()
        for _ in range(1 + a.reps):
            out = ctx.pipeline(model, Cd)
        ctx.sync()
        assert (out[2] == 0).all()
    print("profile target done")


if __name__ == "__main__":
    main()
