"""Generates tests/golden/round2_golden.npz (run once, in the authoring container; nothing here runs on the GPU box).

Imports the REFERENCE's own Python (read-only, /root/reference) and records, on small seeded instances of every family:
  * solvers.dual_computation.compute_oracle_duals(C, noise)        -> (u*, v*) f64[n]  for noise 0 and 1e-3
    (SciPy matching -> difference-constraint Bellman-Ford -> gauge fix -> np.random.seed(42) noise; :13-115)
  * gnn.features.compute_row_features_torch(torch.from_numpy(C))   -> feat f32[n,21]   (the torch variant, :246-351)
The matrices come from the repo's generators and are stored so the fixture is self-contained.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
from gnn.features import compute_row_features_torch  # noqa: E402  (reference)
# the reference's solvers/__init__ imports its compiled `lap`; dual_computation.py itself needs numpy + scipy only
_ds = importlib.util.spec_from_file_location("ref_dual_computation", "/root/reference/solvers/dual_computation.py")
_dm = importlib.util.module_from_spec(_ds)
_ds.loader.exec_module(_dm)
compute_oracle_duals = _dm.compute_oracle_duals  # (reference)

_spec = importlib.util.spec_from_file_location(
    "b200_generators", os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200", "solvers", "generators.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

out = {}
for fam, n in (("uniform", 48), ("sparse", 40), ("sparse1e6", 32), ("metric", 40), ("clustered", 36)):
    C = gen.make_instance(fam, n, seed=7)
    out[f"duals/{fam}/C"] = C
    for tag, noise in (("0", 0.0), ("1e-3", 1e-3)):
        u, v = compute_oracle_duals(C.copy(), noise_level=noise)
        out[f"duals/{fam}/u_{tag}"], out[f"duals/{fam}/v_{tag}"] = u, v
for fam, n in (("uniform", 96), ("sparse", 64), ("sparse1e6", 64), ("metric", 80), ("clustered", 72)):
    C = gen.make_instance(fam, n, seed=11)
    out[f"tfeat/{fam}/C"] = C
    out[f"tfeat/{fam}/feat"] = compute_row_features_torch(torch.from_numpy(C)).numpy()
np.savez_compressed(os.path.join(HERE, "round2_golden.npz"), **out)
print("wrote", len(out), "arrays")
