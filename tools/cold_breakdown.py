"""Cycle breakdown of the COLD solve (lapjv_internal: column reduction, 2 x ARR, augmentation) on the profile build.
Usage: B200LAP_PROFILE_LIB=1 python tools/cold_breakdown.py [n ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lap
from b200lap import trace_dict
from solvers import generators as gen
ctx = b200lap.default_context(0)
sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192]
for n in sizes:
    for fam in ("uniform", "metric", "sparse", "clustered"):
        C = gen.make_instance(fam, n, 42)
        Cd = torch.from_numpy(C.astype(np.float32)).cuda()
        ctx.solve_cold(Cd); ctx.sync()
        t0 = time.perf_counter(); out = ctx.solve_cold(Cd, want_trace=True); ctx.sync(); ms = (time.perf_counter() - t0) * 1e3
        d = trace_dict(out[3][0].cpu().numpy())
        tot = max(d["cyc_total"], 1)
        other = tot - d["cyc_relax"] - d["cyc_collect"] - d["cyc_arr_scan"] - d["cyc_arr_serial"]
        print(f"{fam:9s} n={n}: {ms:8.1f} ms | free after col-red {d['free_after_cr']} | ARR {d['arr_iters']} iters: scan {100*d['cyc_arr_scan']/tot:4.1f}% ({d['cyc_arr_scan']/max(d['arr_iters'],1):6.0f} cyc/iter) "
              f"serial {100*d['cyc_arr_serial']/tot:4.1f}% ({d['cyc_arr_serial']/max(d['arr_iters'],1):6.0f}) | paths {d['aug_paths']}: relax {d['relax_cols']} steps {100*d['cyc_relax']/tot:4.1f}% "
              f"({d['cyc_relax']/max(d['relax_cols'],1):6.0f} cyc/step) collect {d['collect_calls']} {100*d['cyc_collect']/tot:4.1f}% ({d['cyc_collect']/max(d['collect_calls'],1):6.0f}) | other {100*other/tot:4.1f}%", flush=True)
