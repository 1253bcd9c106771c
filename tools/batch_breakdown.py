"""Per-instance SM-cycle breakdown of the batched seeded solve.  Needs the measurement build of the library:
    python gnn-accelerated-lap-warm-start-pipeline_b200/build.py --profile
    B200LAP_PROFILE_LIB=1 python tools/batch_breakdown.py [regpath]
Register-resident path: `row` = cycles from the top of a relax step until the scanned row has arrived, `bar` = cycles
in the step's barrier (trace words 13, 14); the rest of a step is arithmetic, hit publication and bookkeeping."""
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
for reg in ([int(sys.argv[1])] if len(sys.argv) > 1 else [1, 0]):
    ctx.set_option("solver_regpath", reg)
    out = ctx.solve_seeded(Cd, u64, v64, want_trace=True); ctx.sync()
    tr = out[3].cpu().numpy()
    order = np.argsort(-tr[:, 15])
    print(f"regpath={reg}")
    print("idx family      total_Mcyc relax_steps relax_Mcyc cyc/step  row/step  bar/step collects collect_Mcyc cyc/collect replay/coll rec/coll paths  other_Mcyc")
    for i in list(order[:6]) + list(order[20:22]) + list(order[40:42]):
        t = tr[i]
        other = t[15] - t[11] - t[12]
        rs, cs = max(1, t[9]), max(1, t[8])
        print(f"{i:3d} {batch[i][0]:10s} {t[15]/1e6:10.1f} {t[9]:11d} {t[11]/1e6:10.1f} {t[11]/rs:8.0f} {t[13]/rs:9.0f} {t[14]/rs:9.0f} {t[8]:8d} "
              f"{t[12]/1e6:12.1f} {t[12]/cs:11.0f} {t[17]/cs:11.0f} {t[16]/cs:8.1f} {t[7]:5d} {other/1e6:10.1f}")
    if reg == 1:
        i = order[0]
        t = tr[i]
        rs = max(1, t[9])
        names = ["addr(LDS chain)", "issue+swap", "row arrives(slack)", "compute 4 cols", "hit publish", "barrier", "nh read", "post"]
        cs = max(1, t[8])
        print(f"  SCAN depth at step start: 1:{t[40]} 2:{t[41]} 3-4:{t[42]} 5+:{t[43]}   hits per step: 0:{t[44]} 1:{t[45]} 2+:{t[46]} (hits in those: {t[47]})")
        print(f"  collect (thread 0): transpose+bar={t[28]/cs:.0f} scan+bar={t[29]/cs:.0f} flags+bar={t[38]/cs:.0f} replay+bar={t[39]/cs:.0f}")
        for o, who in ((0, "thread 0"), (1, "thread 33")):
            print(f"  {who}: " + "  ".join(f"{names[k]}={t[20 + 10 * o + k] / rs:.0f}" for k in range(8)))
