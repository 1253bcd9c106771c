"""Cycle breakdown of one large seeded solve per cluster size (profile build: B200LAP_PROFILE=1 python build.py --force)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from bench import named_state_dict
ctx = b200lap.default_context(0)
model = b200lap.Model(ctx, named_state_dict(), topk=16)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device="cuda").manual_seed(42)
Cd = torch.rand((1, n, n), generator=g, device="cuda", dtype=torch.float32)
u64, v64, _ = ctx.predict_duals(model, Cd)
for nc in (8,):
    ctx.set_option("solver_cluster", nc)
    out = ctx.solve_seeded(Cd, u64, v64, want_trace=True); ctx.sync()
    t = out[3].cpu().numpy()[0]
    other = t[15] - t[11] - t[12] - t[13] - t[14]
    print(f"n={n} cluster={nc}: total {t[15]/1e6:9.1f} Mcyc | relax {t[9]} steps {t[11]/1e6:9.1f} Mcyc ({t[11]/max(1,t[9]):7.0f}/step) | collect {t[8]} calls {t[12]/1e6:9.1f} Mcyc ({t[12]/max(1,t[8]):8.0f}/call, replay {t[17]/max(1,t[8]):7.0f}, records {t[16]/max(1,t[8]):5.1f}) | paths {t[7]} other {other/1e6:9.1f} Mcyc ({other/max(1,t[7]):8.0f}/path)", flush=True)
