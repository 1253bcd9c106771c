"""Generates tests/golden/feature_rows_golden.npz (run once, in the authoring container; nothing here runs on the GPU box).

The reference's own gnn.features.compute_row_features (imported read-only from /root/reference) evaluated on FULL-SIZE
instances -- n = 2048 of every benchmark family, n = 8192 clustered, n = 16384 uniform: the sizes at which the device
kernels' sampled brackets, list selections and multi-warp groups actually run -- keeping 64 rows of every result.  The
instances are not stored: the tests rebuild them from (family, n, seed) with the repo's generators (seeded NumPy laws).
"""
import importlib.util
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
from gnn.features import compute_row_features  # noqa: E402  (reference)

_spec = importlib.util.spec_from_file_location(
    "b200_generators", os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200", "solvers", "generators.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

CASES = [("uniform", 2048, 101), ("sparse", 2048, 102), ("metric", 2048, 103), ("clustered", 2048, 104), ("sparse1e6", 2048, 105),
         ("clustered", 8192, 106), ("uniform", 16384, 107)]
out = {}
for fam, n, seed in CASES:
    t0 = time.time()
    C = gen.make_instance(fam, n, seed=seed)
    feat = compute_row_features(C)
    rows = np.unique(np.concatenate([np.arange(16), np.arange(n // 2 - 8, n // 2 + 8), np.arange(n - 16, n),
                                     np.random.default_rng(seed).integers(0, n, 16)]))
    out[f"{fam}/{n}/{seed}/rows"] = rows.astype(np.int32)
    out[f"{fam}/{n}/{seed}/feat"] = feat[rows]
    print(fam, n, seed, f"{time.time() - t0:.1f} s", flush=True)
    del C, feat
np.savez_compressed(os.path.join(HERE, "feature_rows_golden.npz"), **out)
