"""DatasetLoader (b200lap/datasets.py) against the reference's loader contract (scripts/gnn_benchmark.py:295-365):
directory layout by size, filtering on `n`, the (C, u_true, v_true) tuples, max_instances_per_size, missing files."""
import numpy as np
import pytest

from b200lap.datasets import DatasetLoader, save_npz
from solvers import generators as gen


def _write(tmp_path, split, sizes):
    d = tmp_path / "generated/processed" / split / "full"
    d.mkdir(parents=True)
    items = []
    for k, n in enumerate(sizes):
        C = gen.make_instance("uniform", n, seed=k)
        items.append((C, np.arange(n, dtype=np.float64) + k, -np.arange(n, dtype=np.float64)))
    save_npz(d / "test.npz", items)
    return items


def test_loader_filters_by_size_and_limits(tmp_path, capsys):
    items = _write(tmp_path, "small", [8, 12, 8, 8, 12])
    got = DatasetLoader(str(tmp_path)).load_instances([8, 12, 2048], max_instances_per_size=2)
    assert sorted(got) == [8, 12, 2048]
    assert len(got[8]) == 2 and len(got[12]) == 2 and got[2048] == []
    C, u, v = got[8][1]
    assert C.shape == (8, 8) and C.dtype == np.float64
    assert np.array_equal(C, items[2][0]) and np.array_equal(u, items[2][1]) and np.array_equal(v, items[2][2])
    assert np.array_equal(got[12][0][0], items[1][0])
    assert "Test file not found" in capsys.readouterr().out       # the mid_2048 split does not exist


def test_loader_needs_h5py_for_hdf5(tmp_path):
    d = tmp_path / "generated/processed/small/full"
    d.mkdir(parents=True)
    (d / "test.h5").write_bytes(b"\x89HDF\r\n\x1a\n")
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py is installed here")
    except ImportError:
        pass
    with pytest.raises(ImportError, match="h5py"):
        DatasetLoader(str(tmp_path)).load_instances([512])


def test_sizes_beyond_the_datasets_are_reported(tmp_path, capsys):
    assert DatasetLoader(str(tmp_path)).load_instances([8192]) == {8192: []}
    assert "No dataset found" in capsys.readouterr().out
