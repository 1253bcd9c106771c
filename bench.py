#!/usr/bin/env python
"""bench.py -- LAP instances/s of the warm-start hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (N=1): BASELINE.json configs[1] -- a batch of 64 mixed-family (uniform/sparse/metric/clustered)
n=2048 instances, random-init OneGNN (hidden 192, layers 4, k 16, torch.manual_seed(0)), synthetic data
snapped to the binary32 grid.  One step = features -> OneGNN -> min-trick -> seeded JV over the whole
batch.  N>1: one process per GPU (torchrun), every rank owns its own 64 instances (weak scaling, no
data-path collective); the timed region is bracketed by barrier + synchronize, time = max over ranks.

  value      device-resident throughput (C already in HBM as binary32), CUDA events on the context stream
  e2e        the same step through the host-buffer C ABI (b200lap_pipeline_batch): pinned host float64 in,
             host int64 assignments out, host<->device copies inside the timed region
  roofline   dominant dense-pass kernel (the row-feature sweep: k_row_features_group + the redo pass of
             k_row_features_smem) at algorithmic bytes = one read of C = 4 n^2 B per instance, timed live with CUDA
             events; `traffic` = DRAM bytes of one launch from the committed ncu capture of the same kernel
             (profiles/r02_ncu_row_features_group.json); per-kernel tables in `dense_pass_*`
  other_configs   BASELINE configs 1, 3, 4, 5 (reduced batch for 5), timed in the same run on rank 0 at N=1
  cpu_baseline   the oracle's CPU pipeline (NumPy features, NumPy OneGNN, reference-compiled lapjv_seeded when
             oracle/_ref travelled, else the C port) on a bounded sample, on this box's host cores; plus SciPy.

--impl reference times that CPU pipeline as the headline line instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

# One BLAS/OpenMP thread per process, as the reference's scripts pin them (scripts/gnn_benchmark.py:25-31).  This
# must happen BEFORE numpy is imported anywhere in the process tree: worker processes are spawned, re-import this
# module and inherit the environment.
for _var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ[_var] = "1"

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200")
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")    # the reference's own gnn/ + lap/, built by oracle/Makefile (pyref)
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

N_INST = 2048
BATCH = 64
FAMILIES = ("uniform", "sparse", "metric", "clustered")
METRIC = "LAP instances/s end-to-end (features -> OneGNN -> min-trick -> lapjv_seeded), mixed-family n=2048"


def load_traffic(workload: str):
    """DRAM read+write bytes of one k_row_features_group launch on `workload`, from the committed ncu --set full capture
    of the kernel this bench times (profiles/r02_ncu_row_features_group.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_row_features_group.json")) as f:
            rec = json.load(f)[workload]
        return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"])
    except Exception:  # noqa: BLE001
        return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def bench_config(world: int) -> dict:
    """The `config` object of both arms (the driver compares them key by key)."""
    return {"workload": f"{BATCH} mixed-family n={N_INST} instances per GPU (BASELINE configs[1], mid2048), random-init OneGNN h192/L4/k16",
            "families": list(FAMILIES), "storage": "binary32 C on the device (exact), binary64 solver arithmetic; binary64 on the host",
            "l2": "inputs (1.07 GB/step) larger than L2", "parallelism": f"instance-sharded x{world}", "steps_in_flight": int(os.environ.get("B200LAP_LANES", "8"))}


def named_state_dict():
    import torch
    from gnn.one_gnn import OneGNN
    torch.manual_seed(0)
    module = OneGNN(21, hidden=192, layers=4, dropout=0.1, topk=16).eval()
    return {k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def make_batch(rank: int, batch: int = BATCH, n: int = N_INST, out=None):
    """The rank's instances as float64 [batch, n, n], written into `out` when given (a pinned buffer: one host
    copy per rank keeps eight ranks of one box inside its RAM)."""
    from solvers import generators as gen
    if out is None:
        out = np.empty((batch, n, n), dtype=np.float64)
    fams = []
    for k in range(batch):
        fam = FAMILIES[k % len(FAMILIES)]
        out[k] = gen.make_instance(fam, n, seed=42 + rank * batch + k)
        assert np.array_equal(out[k].astype(np.float32).astype(np.float64), out[k]), "instances must be binary32-representable"
        fams.append(fam)
    return out, fams


# ---- CPU pipeline (the reference arm / cpu_baseline) ---------------------------------------------------
# kind "reference": the reference's OWN Python pipeline, unmodified, from oracle/_ref/pyref (gnn.compute_row_features,
# torch-CPU gnn.OneGNN, lap.lapjv_seeded = its Cython binding over its C++), glued exactly as
# scripts/gnn_benchmark.py:226-262,289 does on a CPU device.  kind "port": the oracle's restatement (NumPy features,
# NumPy OneGNN, C port of the solver) -- only when oracle/_ref/pyref did not travel.
_W = {"sd": None, "kind": None, "C": None, "model": None, "torch": None}


def pyref_available() -> bool:
    return os.path.isdir(os.path.join(PYREF, "lap")) and os.path.isdir(os.path.join(PYREF, "gnn"))


def _cpu_init(sd, kind, path, shape):
    _W["sd"], _W["kind"] = sd, kind
    _W["C"] = np.load(path, mmap_mode="c") if path else None
    assert _W["C"] is None or tuple(_W["C"].shape) == tuple(shape)
    if kind == "reference":
        sys.path.insert(0, PYREF)             # ahead of this repo's own drop-in `gnn` / `lap` packages
        import torch
        import gnn as ref_gnn
        import lap as ref_lap
        assert os.path.dirname(ref_gnn.__file__).startswith(PYREF) and os.path.dirname(ref_lap.__file__).startswith(PYREF)
        torch.set_num_threads(1)
        model = ref_gnn.OneGNN(21, hidden=192, layers=4, dropout=0.1, topk=16)
        model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
        _W["model"], _W["torch"], _W["feat"], _W["lap"] = model.eval(), torch, ref_gnn.compute_row_features, ref_lap
    else:
        import oracle
        oracle.port_lib()


def _cpu_threads(_):
    """What the worker's numeric libraries actually run with (must be 1 each)."""
    from threadpoolctl import threadpool_info
    info = [(d.get("internal_api"), d.get("num_threads")) for d in threadpool_info()]
    if _W["torch"] is not None:
        info.append(("torch", _W["torch"].get_num_threads()))
    return info


def _cpu_worker(k):
    C = np.ascontiguousarray(_W["C"][k])      # a view of the page-cached input file, float64 [n, n]
    n = C.shape[0]
    t0 = time.perf_counter()
    if _W["kind"] == "reference":
        torch = _W["torch"]
        C = np.asarray(C, dtype=np.float64)                                   # scripts/gnn_benchmark.py:226
        with torch.inference_mode():
            row_feat = _W["feat"](C)                                          # :242
            row_tensor = torch.from_numpy(row_feat).float().unsqueeze(0)      # :243
            cost_batch = torch.from_numpy(C).float().unsqueeze(0)             # :244
            mask = torch.ones((1, n), dtype=torch.bool)                       # :246
            outputs = _W["model"](row_tensor, cost=cost_batch, mask=mask)     # :248
            u_pred = outputs["u"].squeeze(0)[:n].cpu().numpy()                # :251,261
        v_pred = np.min(C - u_pred[:, None], axis=0)                          # :262
        _W["lap"].lapjv_seeded(C, u_pred.astype(np.float64), v_pred.astype(np.float64))   # :289 + solvers/lap_solver.py:81-101
    else:
        from oracle import pipeline_np
        pipeline_np.solve(C, _W["sd"], topk=16, use_ref=False)
    return time.perf_counter() - t0


def _scipy_worker(k):
    from scipy.optimize import linear_sum_assignment
    C = np.ascontiguousarray(_W["C"][k])
    t0 = time.perf_counter()
    linear_sum_assignment(C)
    return time.perf_counter() - t0


class CpuArm:
    """`procs` worker processes (spawned: they inherit the 1-thread environment set at the top of this file) over
    the instances of one float64 [batch, n, n] file, handed out one at a time (dynamic balance)."""

    def __init__(self, Ch: np.ndarray, procs: int):
        import multiprocessing as mp
        import tempfile
        import oracle
        self.kind = "reference" if pyref_available() else "port"
        if self.kind == "port":
            oracle.build()
        self.procs, self.batch = procs, int(Ch.shape[0])
        fd, self.path = tempfile.mkstemp(suffix=".npy", prefix="b200lap_bench_", dir="/dev/shm" if _shm_has(Ch.nbytes) else None)
        os.close(fd)
        np.save(self.path, Ch)
        self.pool = mp.get_context("spawn").Pool(procs, initializer=_cpu_init, initargs=(named_state_dict(), self.kind, self.path, Ch.shape))
        info = self.pool.map(_cpu_threads, range(procs), chunksize=1)
        self.threads = sorted({t for w in info for t in w})
        assert all(nt == 1 for _, nt in self.threads), f"worker threads are not pinned to 1: {self.threads}"

    def run(self, fn, count: int):
        t0 = time.perf_counter()
        per = self.pool.map(fn, range(count), chunksize=1)
        return time.perf_counter() - t0, per

    def close(self):
        self.pool.close()
        self.pool.join()
        try:
            os.unlink(self.path)
        except OSError:
            pass


def _shm_has(nbytes: int) -> bool:
    try:
        st = os.statvfs("/dev/shm")
        return st.f_bavail * st.f_frsize > nbytes + (256 << 20)
    except OSError:
        return False


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---- clocks --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of GPU `index`, sampled during the timed region.  NVML in-process (nvidia_ml_py):
    with one rank per GPU, eight ranks forking `nvidia-smi` five times a second contend for the driver's locks with
    the kernel launches they are supposed to observe; the subprocess form is only the fall-back."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys))
        except Exception:  # noqa: BLE001
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            nv, h = self.nvml
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:  # noqa: BLE001
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            bits = (getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4))
            return [str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for b in bits]
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return [c.strip() for c in out.split(",")] if out else None

    def run(self):
        while not self.stop.is_set():
            try:
                row = self.sample()
                if row:
                    self.rows.append(row)
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(0.1 if self.nvml is not None else 0.5)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=3)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [nm for k, nm in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---- the other BASELINE configs (rank 0, N=1): parity-test shapes, timed in the same run -------------------------
def other_configs(ctx, model, args):
    """configs[0] single uniform n=512 (host API latency, with the reference's CPU pipeline beside it), configs[2]
    32 x metric n=4096 (fallback path), configs[3] single uniform n=16384 end to end, configs[4] n=8192 mixed, solver
    only, oracle duals + N(0, sigma) noise (scripts/main_benchmark.py:45, solvers/dual_computation.py:77-115) at
    --batch5 instances."""
    import torch
    import b200lap
    import lap
    from solvers import generators as gen
    ctx.set_overlap(False)       # one batch at a time from here on (the host pipeline of the e2e leg left the lanes on)
    ctx.sync()
    stream = ctx.torch_stream()
    out = {}

    def ev_time(fn, reps=3, warm=1):
        for _ in range(warm):
            fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            r = fn()
        e1.record(stream)
        ctx.sync()
        return e0.elapsed_time(e1) / reps, r

    def counters(tr):
        tr = tr.cpu().numpy()
        return {"took_fallback": int(tr[:, 3].sum()), "aug_paths": int(tr[:, 7].sum()), "relax_cols": int(tr[:, 9].sum()),
                "arr_iters": int(tr[:, 6].sum()), "proj_triggers": int(tr[:, 0].sum())}

    try:
        # config 1: the reference-facing call, lap.lapjv_seeded(C, u, v) after GNNPredictor.predict(C), host buffers
        C = gen.make_instance("uniform", 512, 42)
        pred = b200lap.GNNPredictor(named_state_dict(), device=ctx.device, ctx=ctx)

        def host():
            u, v = pred.predict(C)
            return lap.lapjv_seeded(C, u, v)
        host()
        t0 = time.perf_counter()
        for _ in range(10):
            host()
        ms = (time.perf_counter() - t0) / 10 * 1e3
        dms, r = ev_time(lambda: ctx.pipeline(model, pred.to_device(C), want_trace=True), reps=10)
        out["config1_uniform_512"] = {"host_api_ms": round(ms, 3), "device_resident_ms": round(dms, 3), "counters": counters(r[5])}
    except Exception as exc:  # noqa: BLE001
        out["config1_uniform_512"] = {"error": repr(exc)}
    try:
        n3, B3 = 4096, 32
        Cd = torch.empty((B3, n3, n3), dtype=torch.float32, device="cuda")
        for k in range(B3):
            Cd[k] = torch.from_numpy(gen.make_instance("metric", n3, 42 + k).astype(np.float32)).cuda()
        ms, r = ev_time(lambda: ctx.pipeline(model, Cd, want_trace=True))
        pms, _ = ev_time(lambda: ctx.predict_duals(model, Cd), reps=5, warm=2)
        assert (r[2] == 0).all()
        out["config3_metric_4096_x32"] = {"pipeline_ms": round(ms, 3), "predict_ms": round(pms, 3), "inst_per_s": round(B3 / ms * 1e3, 1),
                                          "counters": counters(r[5])}
        del Cd
    except Exception as exc:  # noqa: BLE001
        out["config3_metric_4096_x32"] = {"error": repr(exc)}
    try:
        n4 = 16384
        Cd = torch.rand((n4, n4), generator=torch.Generator(device="cuda").manual_seed(42), device="cuda", dtype=torch.float32)
        pms, duals = ev_time(lambda: ctx.predict_duals(model, Cd), reps=5)
        t0 = time.perf_counter()
        x, y, rc, tr = ctx.solve_seeded(Cd, duals[0], duals[1], want_trace=True)
        ctx.sync()
        sms = (time.perf_counter() - t0) * 1e3
        assert int(rc[0]) == 0 and bool((torch.sort(x[0].long()).values == torch.arange(n4, device="cuda")).all())
        out["config4_uniform_16384"] = {"predict_ms": round(pms, 3), "solve_ms": round(sms, 1), "inst_per_s": round(1e3 / (pms + sms), 4),
                                        "counters": counters(tr)}
        del Cd
    except Exception as exc:  # noqa: BLE001
        out["config4_uniform_16384"] = {"error": repr(exc)}
    try:
        n5, B5 = 8192, int(args.batch5)
        Cd = torch.empty((B5, n5, n5), dtype=torch.float32, device="cuda")
        for k in range(B5):
            Cd[k] = torch.from_numpy(gen.make_instance(FAMILIES[k % 4], n5, 42 + k).astype(np.float32)).cuda()
        t0 = time.perf_counter()
        xc, yc, rcc, vfin = ctx.solve_cold(Cd, want_v=True)
        ctx.sync()
        cold_ms = (time.perf_counter() - t0) * 1e3
        u_opt = (Cd.double() - vfin[:, None, :]).min(dim=2).values
        rec = {"batch": B5, "oracle_duals_by_cold_solve_ms": round(cold_ms, 1)}
        for sigma in (0.0, 1e-3, 1e-2):
            gn = torch.Generator(device="cuda").manual_seed(42)
            u = u_opt + sigma * torch.randn(u_opt.shape, generator=gn, device="cuda", dtype=torch.float64)
            v = vfin + sigma * torch.randn(u_opt.shape, generator=gn, device="cuda", dtype=torch.float64)
            t0 = time.perf_counter()
            x, y, rc, tr = ctx.solve_seeded(Cd, u, v, want_trace=True)
            ctx.sync()
            ms = (time.perf_counter() - t0) * 1e3
            assert (rc == 0).all()
            rec[f"sigma_{sigma:g}"] = {"solve_ms": round(ms, 1), "inst_per_s": round(B5 / ms * 1e3, 3), "counters": counters(tr)}
        out["config5_mixed_8192_oracle_duals_plus_noise"] = rec
        del Cd
    except Exception as exc:  # noqa: BLE001
        out["config5_mixed_8192_oracle_duals_plus_noise"] = {"error": repr(exc)}
    torch.cuda.empty_cache()
    return out


# ---- the B200 arm ----------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import b200lap

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    # one process per GPU: keep this rank's CPU threads and its pinned host buffers on the GPU's NUMA node
    numa_node = b200lap.bind_to_device_numa_node(local) if world > 1 and os.environ.get("B200LAP_NO_NUMA_BIND", "0") != "1" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = b200lap.default_context(local)
    model = b200lap.Model(ctx, named_state_dict(), topk=16)
    n, B = N_INST, BATCH
    Cp = torch.empty((B, n, n), dtype=torch.float64).pin_memory()     # the e2e leg's host input; also the only host copy
    Ch, fams = make_batch(rank, out=Cp.numpy())
    Cd = torch.empty((B, n, n), dtype=torch.float32, device="cuda")
    for k in range(B):                                                # exact: instances live on the binary32 grid (checked in make_batch)
        Cd[k] = torch.from_numpy(Ch[k].astype(np.float32)).cuda()
    stream = ctx.torch_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -- device-resident throughput: B200LAP_LANES batches in flight (the context's lanes rotate), so the SMs one
    #    64-instance solve leaves idle (one CTA per instance, 148 SMs) work on the neighbouring steps
    def step_resident():
        return ctx.pipeline(model, Cd)

    LANES = int(os.environ.get("B200LAP_LANES", "8"))
    ctx.set_overlap(LANES)
    # N > 1, optional (B200LAP_DYNAMIC_QUEUE=1): the job as world x steps batch-units drained from ONE queue
    # (b200lap.WorkQueue: an atomic counter in the process group's store), unit u = the batch of rank u % world; every rank
    # then holds every rank's batch on its device (one NCCL broadcast each at set-up, outside the timed region) and claims a
    # unit whenever a lane frees up.  Measured at N = 2 (profiles/r02_bench_2gpu_*.json): the ranks' mixed-family batches are
    # statistically alike, both ranks end up with the same number of units and the claim back-pressure costs 5 %
    # (4216 vs 4442 inst/s), so the default stays the static split; the queue is for skewed batches.
    dynamic = world > 1 and os.environ.get("B200LAP_DYNAMIC_QUEUE", "0") == "1"
    batches = [Cd]
    if dynamic:
        batches = []
        for b in range(world):
            t = Cd if b == rank else torch.empty_like(Cd)
            dist.broadcast(t, src=b)
            batches.append(t)
        torch.cuda.synchronize()
    # every lane gets at least one untimed step: a lane's first call allocates its workspace (cudaMalloc synchronises the
    # device), which otherwise lands inside the timed region as a 29 -> 36-45 ms step (tools/lane_jitter.py)
    args.warmup = max(args.warmup, LANES)
    for _ in range(args.warmup):
        out = step_resident()
    ctx.sync()
    barrier()
    launches0 = ctx.launches
    outs = []
    units_done = args.steps
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if dynamic:
            queue = b200lap.WorkQueue(world * args.steps, name="bench_resident")

            def launch(unit):
                outs.append(ctx.pipeline(model, batches[unit % world]))
                del outs[:-LANES]
                return ctx.last_lane_event()
            units_done = len(b200lap.drain_queue(queue, launch, in_flight=LANES))
        else:
            for _ in range(args.steps):
                outs = (outs + [step_resident()])[-LANES:]
        ctx.join()                      # lane 0's stream waits for the other lanes on the device ...
        e1.record(stream)               # ... so this event closes the work of all lanes
        ctx.sync()
        barrier()
        ms_resident = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    ctx.set_overlap(False)
    for o in outs:
        assert (o[2].cpu().numpy() == 0).all(), o[2]
    if not dynamic:
        out = outs[-1]
        for o in outs:
            assert torch.equal(o[0], out[0]) and torch.equal(o[1], out[1]), "the lanes disagree"
    else:
        out = step_resident()          # this rank's own batch, for the host-path comparison below
        ctx.sync()
    t = torch.tensor([ms_resident], dtype=torch.float64, device="cuda")
    ms_by_rank = [round(ms_resident / args.steps, 3)]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        ms_by_rank = [round(float(e.item()) / args.steps, 3) for e in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    tu = torch.tensor([float(units_done)], dtype=torch.float64, device="cuda")
    units_all = [torch.zeros_like(tu) for _ in range(world)]
    if world > 1:
        dist.all_gather(units_all, tu)
    units_per_rank = [int(u.item()) for u in units_all] if world > 1 else [args.steps]
    # one batch at a time (latency of a step when nothing else is in flight)
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record(stream)
    for _ in range(3):
        step_resident()
    eb.record(stream)
    ctx.sync()
    ms_step_alone = ea.elapsed_time(eb) / 3

    # -- per-kernel dense pass (CUDA events on the launching stream), same resident batch
    u = torch.zeros((B, n), dtype=torch.float32, device="cuda")

    def timed(fn, reps):
        fn(); ctx.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        ctx.sync()
        return a.elapsed_time(b) / reps

    reps = max(3, args.steps)
    feat, topv = ctx.row_features(Cd, topk=16)
    u64, v64, u32 = ctx.predict_duals(model, Cd)
    bytes_one_read = 4.0 * n * n * B
    dense = {}
    for name, fn in (("col_argmin", lambda: ctx.col_argmin(Cd)),
                     ("row_features", lambda: ctx.row_features(Cd, topk=16)),
                     ("onegnn_mlp", lambda: ctx.onegnn_forward(model, feat, topv)),
                     ("min_trick", lambda: ctx.min_trick(Cd, u32)),
                     ("front_end", lambda: ctx.front_end(Cd, u64, v64)),
                     ("predict_duals", lambda: ctx.predict_duals(model, Cd))):
        ms = timed(fn, reps)
        dense[name] = {"ms": round(ms, 4), "GBps_one_read_of_C": round(bytes_one_read / (ms * 1e-3) / 1e9, 1)}
    ctx.row_features(Cd, topk=16)
    redo_rows = ctx.feature_redo_rows()
    solve_ms = timed(lambda: ctx.solve_seeded(Cd, u64, v64), max(2, args.steps // 2))
    dense["solve_seeded"] = {"ms": round(solve_ms, 3)}
    # what the random-init seeds make the solver do (the headline depends on it), per family
    counters = None
    try:
        tr = ctx.solve_seeded(Cd, u64, v64, want_trace=True)[3].cpu().numpy()
        counters = {}
        for f in FAMILIES:
            rows = tr[[k for k in range(B) if fams[k] == f]]
            counters[f] = {"took_fallback_rate": round(float(rows[:, 3].mean()), 3), "mean_aug_paths": round(float(rows[:, 7].mean()), 1),
                           "mean_proj_triggers": round(float(rows[:, 0].mean()), 1), "mean_tight_edges": round(float(rows[:, 1].mean()), 1)}
    except Exception as exc:  # noqa: BLE001
        counters = {"error": repr(exc)}

    # -- the n = 16384 dense pass (config 4: 1 GiB binary32 C), the size the HBM target is quoted on
    big = None
    if rank == 0 and not args.skip_big:
        try:
            nb = 16384
            g = torch.Generator(device="cuda").manual_seed(42)
            Cb = torch.rand((nb, nb), generator=g, device="cuda", dtype=torch.float32)
            fb, tb = ctx.row_features(Cb, topk=16)
            ub64, vb64, ub32 = ctx.predict_duals(model, Cb)
            one = 4.0 * nb * nb
            big = {}
            for name, fn in (("col_argmin", lambda: ctx.col_argmin(Cb)),
                             ("row_features", lambda: ctx.row_features(Cb, topk=16)),
                             ("onegnn_mlp", lambda: ctx.onegnn_forward(model, fb, tb)),
                             ("min_trick", lambda: ctx.min_trick(Cb, ub32)),
                             ("front_end", lambda: ctx.front_end(Cb, ub64, vb64)),
                             ("predict_duals", lambda: ctx.predict_duals(model, Cb))):
                ms = timed(fn, 5)
                big[name] = {"ms": round(ms, 4), "GBps_one_read_of_C": round(one / (ms * 1e-3) / 1e9, 1)}
            del Cb, fb, tb
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001
            big = {"error": repr(exc)}

    # -- end to end through the host-buffer C ABI (pinned host float64 in, host int64 out), two batches in flight:
    #    b200lap_pipeline_batch_submit / _wait; every step uploads its 2 GiB of binary64 matrices and downloads its
    #    assignments inside the timed region, the upload of step k+1 overlaps the solve of step k
    torch.cuda.set_device(local)          # the host entry points use the library's default context: make it this rank's device
    # host-side marshalling of the upload (csrc/host_narrow.cpp): threads per rank scaled to the ranks sharing the host
    lws = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    # measured (profiles/r02_host_narrow_sweep.txt): one rank gains 35 % from narrowing 70 % of the batch on 12 host threads; the
    # box's 16 cores narrow at most ~85 GB/s in total, so the share shrinks with the ranks per host: +3 % at two ranks (50 % on 8
    # threads each), +9 % at four (50 % on 4), and at eight the plain binary64 DMA wins (5258 vs 4342 inst/s with 30 % on 2)
    hn_threads, hn_percent = {1: (12, 70), 2: (8, 50), 4: (4, 50)}.get(lws, (1, 0))
    os.environ.setdefault("B200LAP_HOST_NARROW_THREADS", str(max(1, min(hn_threads, host_cores() // lws))))
    os.environ.setdefault("B200LAP_HOST_NARROW_PERCENT", str(hn_percent))
    hn_t, hn_p = ctypes.c_int(0), ctypes.c_int(0)
    h2d_bytes = int(ctx.lib.b200lap_host_narrow_config(B, n, ctypes.byref(hn_t), ctypes.byref(hn_p)))
    hp = b200lap.HostPipeline(named_state_dict(), topk=16, lanes=LANES)
    hbuf = [(torch.empty((B, n), dtype=torch.int64).pin_memory(), torch.empty((B, n), dtype=torch.int64).pin_memory(),
             np.zeros(B, dtype=np.int32)) for _ in range(LANES)]

    def run_e2e(steps):
        pending = []
        for s_ in range(steps):
            xh_, yh_, rc_ = hbuf[s_ % LANES]
            pending.append(hp.submit(Cp, xh_, yh_, rc_))
            if len(pending) == LANES:
                hp.wait(pending.pop(0))
        while pending:
            hp.wait(pending.pop(0))

    e2e_steps = max(LANES, args.steps)
    run_e2e(LANES)
    barrier()
    t0 = time.perf_counter()
    run_e2e(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    # the synchronous single-call form of the same path (b200lap_pipeline_batch), one batch at a time
    xh, yh, rch = hbuf[0]
    t0 = time.perf_counter()
    hp.run(Cp, xh, yh, rch)
    e2e_sync_s = time.perf_counter() - t0
    assert (rch == 0).all() and (hbuf[1][2] == 0).all()
    # resident, asynchronous and synchronous host paths must agree bit for bit
    assert np.array_equal(out[0].cpu().numpy().astype(np.int64), xh.numpy()), "host and resident paths disagree"
    assert np.array_equal(hbuf[1][0].numpy(), xh.numpy()) and np.array_equal(hbuf[1][1].numpy(), yh.numpy()), "asynchronous host path disagrees"

    others = None
    if rank == 0 and world == 1 and not args.skip_configs:
        others = other_configs(ctx, model, args)
    if rank == 0:
        peak, which = load_peaks()
        fk = dense["row_features"]
        achieved = fk["GBps_one_read_of_C"]
        cpu = None
        if not args.skip_cpu and world == 1:      # the CPU sample is timed on rank 0 at N = 1 only
            procs = host_cores()
            sample = min(B, max(procs, 32))
            arm = CpuArm(Ch[:sample], procs)
            try:
                arm.run(_cpu_worker, min(sample, procs))                       # untimed: page in torch / BLAS / the solver
                wall, per = arm.run(_cpu_worker, sample)
                swall, sper = arm.run(_scipy_worker, sample)
            finally:
                arm.close()
            cpu = {"value": round(sample / wall, 3), "unit": "instances/s", "cores": procs, "kind": arm.kind,
                   "sample": f"the first {sample} of the step's {B} mixed-family n={N_INST} instances, handed one at a time to {procs} "
                             f"worker processes, 1 thread each (mean latency {float(np.mean(per)):.2f} s)",
                   "pipeline": "reference's own gnn.compute_row_features + torch-CPU OneGNN + lap.lapjv_seeded (oracle/_ref/pyref)"
                               if arm.kind == "reference" else "oracle restatement (NumPy features/MLP + C port of the solver)",
                   "worker_threads": [list(t) for t in arm.threads],
                   "scipy_linear_sum_assignment": {"value": round(sample / swall, 3), "unit": "instances/s",
                                                   "mean_latency_s": round(float(np.mean(sper)), 3)}}
        line = {
            "metric": METRIC, "value": round(world * B / (ms_step * 1e-3), 2), "unit": "instances/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "ms_per_step_by_rank": ms_by_rank, "ms_per_step_alone": round(ms_step_alone, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(world),
            "numa_node_of_rank0": numa_node,
            "scheduling": ("dynamic: world x steps batch-units drained from one queue (b200lap.WorkQueue), units per rank " + str(units_per_rank))
                          if dynamic else "static: every rank steps over its own batch",
            "e2e": {"value": round(world * B / e2e_s, 2), "unit": "instances/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": int(B * n * 8 * 2 + rch.nbytes + 4), "steps": e2e_steps,
                    "host_input_bytes_per_step": int(Ch.nbytes),
                    "upload": (f"first {hn_p.value}% of the instances narrowed to binary32 by {hn_t.value} host threads (exactness-checked), the rest "
                               "uploaded as binary64 by DMA meanwhile and narrowed on the device") if hn_p.value > 0 and hn_t.value > 0 else
                              "whole batch uploaded as binary64 by DMA and narrowed on the device (several ranks share this host's cores and memory)",
                    "api": f"b200lap_pipeline_batch_submit/_wait, {LANES} batches in flight",
                    "one_batch_at_a_time": {"value": round(world * B / e2e_sync_s, 2), "unit": "instances/s", "api": "b200lap_pipeline_batch"}},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "kernel": "row-feature sweep: k_row_features_group<1,64> + the redo pass of k_row_features_smem (21-D features + top-16, one read of C)", "achieved": achieved,
                         "peak": peak, "peak_source": which, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "algorithmic_bytes_per_launch": bytes_one_read, "traffic": load_traffic("n2048_b64"),
                         "rows_handed_to_fallback_kernel": redo_rows,
                         "n16384": None if not big or "row_features" not in big else
                         {"achieved": big["row_features"]["GBps_one_read_of_C"], "frac": round(big["row_features"]["GBps_one_read_of_C"] / peak, 4),
                          "dense_pass_frac": round(big["predict_duals"]["GBps_one_read_of_C"] / peak, 4), "traffic": load_traffic("n16384_b1")}},
            "dense_pass_n2048_b64": dense,
            "dense_pass_n16384": big,
            "solver_counters": counters,
            "other_configs": others,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---- the reference arm -------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores: every step is the SAME 64
    instances the B200 arm steps over (rank 0's batch), spread over all host cores, one instance per worker at a time."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    procs = host_cores()
    Ch, fams = make_batch(0)
    arm = CpuArm(Ch, procs)
    try:
        for _ in range(args.warmup):
            arm.run(_cpu_worker, BATCH)
        walls = [arm.run(_cpu_worker, BATCH)[0] for _ in range(args.steps)]
    finally:
        arm.close()
    v = BATCH * len(walls) / sum(walls)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": "instances/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1e3 * sum(walls) / len(walls), 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(world),          # the B200 arm's config key for key (the driver compares them); the CPU work is the same at every N
            "cpu_baseline": {"value": round(v, 3), "unit": "instances/s", "cores": procs, "kind": arm.kind,
                             "sample": f"each step = the {BATCH} mixed-family n={N_INST} instances of rank 0's batch over {procs} worker processes, 1 thread each",
                             "pipeline": "reference's own gnn.compute_row_features + torch-CPU OneGNN + lap.lapjv_seeded (oracle/_ref/pyref)"
                                         if arm.kind == "reference" else "oracle restatement (NumPy features/MLP + C port of the solver)",
                             "worker_threads": [list(t) for t in arm.threads]},
            "e2e": {"value": round(v, 3), "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)      # 8 batches are in flight: a few dozen steps amortise the ramp
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg")
    ap.add_argument("--skip-big", action="store_true", help="omit the n=16384 dense-pass table")
    ap.add_argument("--skip-configs", action="store_true", help="omit BASELINE configs 1, 3, 4, 5 (`other_configs`)")
    ap.add_argument("--batch5", type=int, default=8, help="instances of config 5 (n=8192, oracle duals + noise) on this GPU")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        args.warmup = min(args.warmup, 1)      # each step is tens of CPU-seconds: keep the arm inside a few minutes
        args.steps = min(args.steps, 3)
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
