"""Generates tests/golden/dense_golden.npz (run once, in the authoring container).

Imports the REFERENCE's own Python (read-only, /root/reference) and records, on small
seeded instances of every cost family:
  * gnn.features.compute_row_features(C)                       -> feat  f32[n,21]
  * gnn.one_gnn.OneGNN(...).eval()(feat, cost=C32, mask=ones)  -> u     f32[n]
  * np.min(C - u[:,None], axis=0)  (scripts/gnn_benchmark.py:262) -> v  f64[n]
for (a) a SMALL model (hidden=32, layers=2, k=8) whose full state_dict is stored, and
(b) the NAMED architecture (hidden=192, layers=4, k=16) initialised with
torch.manual_seed(0); its weights are not stored (1.4 MB) -- a checksum of every tensor
is, and the tests rebuild them through the repo's own OneGNN mirror, which must reproduce
the checksum before outputs are compared.
The matrices come from the repo's generators (already checked to follow the reference's
laws) and are stored so the fixture is self-contained.  Nothing here runs on the GPU box.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200"))
sys.path.insert(0, "/root/reference")

from gnn.features import compute_row_features  # noqa: E402  (reference)
from gnn.one_gnn import OneGNN  # noqa: E402  (reference)

# the repo's generators live in a package that is also called "solvers"; load by path
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "b200_generators", os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200", "solvers", "generators.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

FAMILIES = ("uniform", "sparse", "sparse1e6", "metric", "clustered")
SIZES = {"uniform": 96, "sparse": 64, "sparse1e6": 64, "metric": 80, "clustered": 72}


def run_model(model, C):
    feat = compute_row_features(C)
    n = C.shape[0]
    with torch.inference_mode():
        row = torch.from_numpy(feat).float().unsqueeze(0)
        cost = torch.from_numpy(C).float().unsqueeze(0)
        mask = torch.ones((1, n), dtype=torch.bool)
        u = model(row, cost=cost, mask=mask)["u"].squeeze(0)[:n].cpu().numpy()
        # pre-centring head output: the scale the u tolerance is stated against
        h = model.input_proj(row)
        for blk in model.blocks:
            h = blk(h)
        u_pre = model.pre_out(h).squeeze(-1)
        h = h + model._sparse_refine(h, cost, u_pre, mask)
        raw = model.row_out(h).squeeze(-1).squeeze(0).cpu().numpy()
    v = np.min(C - u[:, None], axis=0)
    return feat, u, raw, v


def main():
    out = {}
    torch.manual_seed(1234)
    small = OneGNN(21, hidden=32, layers=2, dropout=0.1, topk=8).eval()
    for k, t in small.state_dict().items():
        out["small_sd/" + k] = t.numpy().copy()
    torch.manual_seed(0)
    named = OneGNN(21, hidden=192, layers=4, dropout=0.1, topk=16).eval()
    for k, t in named.state_dict().items():
        a = t.numpy().astype(np.float64)
        out["named_ck/" + k] = np.array([a.sum(), np.abs(a).sum(), (a * np.arange(1, a.size + 1).reshape(a.shape)).sum()])
    for fam in FAMILIES:
        n = SIZES[fam]
        C = gen.make_instance(fam, n, 42)
        out[f"{fam}/C"] = C
        f, u, raw, v = run_model(small, C)
        out[f"{fam}/feat"] = f
        out[f"{fam}/small_u"] = u
        out[f"{fam}/small_raw"] = raw
        out[f"{fam}/small_v"] = v
        f2, u, raw, v = run_model(named, C)
        assert np.array_equal(f, f2)
        out[f"{fam}/named_u"] = u
        out[f"{fam}/named_raw"] = raw
        out[f"{fam}/named_v"] = v
    # an odd-n instance (median/MAD take a single middle element) and a tiny one (k > n)
    for tag, n in (("odd", 33), ("tiny", 5)):
        C = gen.make_instance("uniform", n, 7)
        out[f"{tag}/C"] = C
        f, u, raw, v = run_model(small, C)
        out[f"{tag}/feat"] = f
        out[f"{tag}/small_u"] = u
        out[f"{tag}/small_raw"] = raw
        out[f"{tag}/small_v"] = v
    path = os.path.join(HERE, "dense_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
