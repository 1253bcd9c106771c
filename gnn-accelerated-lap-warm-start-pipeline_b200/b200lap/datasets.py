"""``DatasetLoader`` -- the reference benchmark's instance reader, same class, method and return shape as
/root/reference/scripts/gnn_benchmark.py:295-365, over the dataset schema of /root/reference/data/generate_dataset.py:49-62
(one record per instance: variable-length float64 ``C`` (row-major n*n), ``u``, ``v``; int32 ``n``; plus ``rows``,
``cols``, ``cost``, ``family``, ``noise_std``, ``tag`` which the loader does not read).

Two containers are read through the same code path:
  * ``*.h5``  -- the reference's HDF5 files, through ``h5py`` when it is importable (it is not part of this image:
                 without it a ``.h5`` file raises ``ImportError`` naming the missing module; nothing is silently skipped);
  * ``*.npz`` -- the same record arrays saved with ``numpy.savez`` (``C`` an object array of flat float64 vectors, or a
                 2-D array when every instance has the same size): what ``save_npz`` below writes, used by the tests and
                 by boxes without h5py.
``to_device_batches`` stacks the loaded instances of one size into the float32 / float64 device batches the B200
pipeline consumes (``b200lap.Context.pipeline``), which is what the reference's benchmark loop does one instance at a time.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List

import numpy as np

# size -> dataset directory below <data_dir>/generated/processed (scripts/gnn_benchmark.py:316-333)
_SPLITS = ((512, "small"), (1024, "small"), (1536, "mid_1536"), (2048, "mid_2048"), (3072, "mid_3072"), (4096, "large_4096"))


class _NpzRecords:
    def __init__(self, path):
        self.z = np.load(path, allow_pickle=True)

    def __getitem__(self, key):
        return self.z[key]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.z.close()


def _open(path: Path):
    if path.suffix == ".npz":
        return _NpzRecords(path)
    try:
        import h5py
    except ImportError as exc:
        raise ImportError(f"{path} is an HDF5 dataset and the 'h5py' module is not installed; "
                          f"convert it with b200lap.datasets.save_npz on a machine that has h5py") from exc
    return h5py.File(path, "r")


def save_npz(path, instances) -> None:
    """Writes [(C, u, v), ...] as a record file this loader reads (the reference's field names)."""
    Cs = np.empty(len(instances), dtype=object)
    us = np.empty(len(instances), dtype=object)
    vs = np.empty(len(instances), dtype=object)
    ns = np.empty(len(instances), dtype=np.int32)
    for k, (C, u, v) in enumerate(instances):
        C = np.asarray(C, dtype=np.float64)
        Cs[k], us[k], vs[k], ns[k] = C.reshape(-1), np.asarray(u, dtype=np.float64), np.asarray(v, dtype=np.float64), C.shape[0]
    np.savez(path, C=Cs, u=us, v=vs, n=ns)


class DatasetLoader:
    """Load test instances from the reference's datasets."""

    def __init__(self, data_dir: str):
        self.data_dir = Path(data_dir)

    def _test_file(self, size: int):
        for limit, name in _SPLITS:
            if size <= limit:
                base = self.data_dir / "generated/processed" / name / "full"
                for cand in (base / "test.h5", base / "test.npz"):
                    if cand.exists():
                        return cand
                return base / "test.h5"
        return None

    def load_instances(self, problem_sizes: List[int], max_instances_per_size: int = 50) -> Dict:
        """Dict mapping size -> list of (C, u_true, v_true) tuples (C float64 [n, n])."""
        instances = {}
        for size in problem_sizes:
            instances[size] = []
            test_file = self._test_file(size)
            if test_file is None:
                print(f"⚠️ No dataset found for size {size}")
                continue
            if not test_file.exists():
                print(f"⚠️ Test file not found: {test_file}")
                continue
            print(f"Loading {size}x{size} instances from: {test_file.name}")
            with _open(test_file) as f:
                ns = np.asarray(f["n"][:] if hasattr(f["n"], "shape") else f["n"])
                size_indices = np.flatnonzero(ns == size)[:max_instances_per_size]
                print(f"  Found {len(size_indices)} instances of size {size}x{size}")
                for i in size_indices:
                    n = int(ns[i])
                    C = np.asarray(f["C"][i], dtype=np.float64).reshape(n, n)
                    instances[size].append((C, np.asarray(f["u"][i], dtype=np.float64), np.asarray(f["v"][i], dtype=np.float64)))
        return instances

    @staticmethod
    def to_device_batches(instances, device: int = 0):
        """size -> (C [B, n, n] CUDA tensor, float32 when every entry survives the round trip else float64,
        u_true [B, n] float64, v_true [B, n] float64)."""
        import torch
        out = {}
        for size, items in instances.items():
            if not items:
                continue
            C = np.stack([c for c, _, _ in items])
            C32 = C.astype(np.float32)
            host = C32 if np.array_equal(C32.astype(np.float64), C) else C
            out[size] = (torch.from_numpy(host).to(f"cuda:{device}"),
                         torch.from_numpy(np.stack([u for _, u, _ in items])).to(f"cuda:{device}"),
                         torch.from_numpy(np.stack([v for _, _, v in items])).to(f"cuda:{device}"))
        return out
