"""Repeats the bench's resident leg (8 batches in flight, 24 steps) several times in one process and prints the step
time of each repetition: python tools/lane_jitter.py [reps].  Used to look at run-to-run spread of the lane overlap
(with and without CUDA_DEVICE_MAX_CONNECTIONS raised)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (path set-up, batch generator)
import numpy as np, torch, b200lap  # noqa: E402,E401

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.cuda.set_device(0)
ctx = b200lap.default_context(0)
model = b200lap.Model(ctx, bench.named_state_dict(), topk=16)
Ch, fams = bench.make_batch(0)
Cd = torch.from_numpy(Ch.astype(np.float32)).cuda()
del Ch
stream = ctx.torch_stream()
LANES = int(os.environ.get("B200LAP_LANES", "8"))
ctx.set_overlap(LANES)
for _ in range(3):
    ctx.pipeline(model, Cd)
ctx.sync()
out = []
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    outs = []
    for _ in range(24):
        outs = (outs + [ctx.pipeline(model, Cd)])[-LANES:]
    ctx.join()
    e1.record(stream)
    ctx.sync()
    out.append(round(e0.elapsed_time(e1) / 24, 2))
print("MAX_CONNECTIONS", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "lanes", LANES, "ms/step", out)
