"""Per-instance cycle breakdown of the batched seeded solve (needs the profile build: B200LAP_PROFILE=1 python build.py --force)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from bench import named_state_dict
from solvers import generators as gen
ctx = b200lap.default_context(0)
model = b200lap.Model(ctx, named_state_dict(), topk=16)
n, B = 2048, 64
batch = gen.mixed_batch(n, B, first_seed=42)
Cd = torch.from_numpy(np.stack([c for _, c in batch]).astype(np.float32)).cuda()
u64, v64, _ = ctx.predict_duals(model, Cd)
out = ctx.solve_seeded(Cd, u64, v64, want_trace=True); ctx.sync()
tr = out[3].cpu().numpy()
order = np.argsort(-tr[:, 15])
print("idx family      total_Mcyc relax_steps relax_Mcyc cyc/step collects collect_Mcyc cyc/collect paths  other_Mcyc")
for i in order[:12]:
    t = tr[i]
    other = t[15] - t[11] - t[12] - t[13] - t[14]
    print(f"{i:3d} {batch[i][0]:10s} {t[15]/1e6:10.1f} {t[9]:11d} {t[11]/1e6:10.1f} {t[11]/max(1,t[9]):8.0f} {t[8]:8d} {t[12]/1e6:12.1f} {t[12]/max(1,t[8]):11.0f} {t[7]:5d} {other/1e6:10.1f}")
