"""Prints the solver's own cycle counters (trace words 11-15) for a few instance shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, time
import b200lap
from b200lap import trace_dict
from solvers import generators as gen
ctx = b200lap.default_context(0)
def show(tag, tr, ms):
    d = trace_dict(tr)
    tot = max(d["cyc_total"], 1)
    ghz = tot / (ms * 1e-3) / 1e9
    s = f"{tag}: {ms:9.1f} ms ({ghz:.2f} GHz)  relax {d['relax_cols']} steps {100*d['cyc_relax']/tot:5.1f}% ({d['cyc_relax']/max(d['relax_cols'],1):7.0f} cyc/step)  " \
        f"collect {d['collect_calls']} {100*d['cyc_collect']/tot:5.1f}% ({d['cyc_collect']/max(d['collect_calls'],1):7.0f} cyc)  " \
        f"records/collect {d['collect_records']/max(d['collect_calls'],1):.1f} replay {d['cyc_collect_replay']/max(d['collect_calls'],1):.0f} cyc | relax hits {d['relax_hits']} replay {d['cyc_relax_replay']/max(d['relax_cols'],1):.0f} cyc/step | arr {d['arr_iters']} scan {100*d['cyc_arr_scan']/tot:5.1f}% ({d['cyc_arr_scan']/max(d['arr_iters'],1):7.0f} cyc) serial {100*d['cyc_arr_serial']/tot:5.1f}% ({d['cyc_arr_serial']/max(d['arr_iters'],1):7.0f} cyc)  paths {d['aug_paths']}"
    print(s, flush=True)
for fam, n in (("sparse", 2048), ("uniform", 2048), ("uniform", 4096)):
    C = gen.make_instance(fam, n, 42)
    Cd = torch.from_numpy(C.astype(np.float32)).cuda()
    rng = np.random.default_rng(1)
    u = rng.normal(0, 0.01, n).astype(np.float32).astype(np.float64)
    v = np.min(C - u[:, None], axis=0)
    for name, fn in (("seeded", lambda: ctx.solve_seeded(Cd, torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda(), want_trace=True)),
                     ("cold", lambda: ctx.solve_cold(Cd, want_trace=True))):
        if n == 8192 and name == "cold" and os.environ.get("SKIP_BIG_COLD"): continue
        fn(); ctx.sync()
        t0 = time.perf_counter(); out = fn(); ctx.sync(); ms = (time.perf_counter() - t0) * 1e3
        show(f"{fam} n={n} {name}", out[3][0].cpu().numpy(), ms)
