"""CPU tests: the product library loads and exports every symbol include/b200lap.h declares (no compute
calls: there is no GPU here), the ctypes table covers the header, and the product refuses to compute
without a device instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200lap.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_ \*]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text, flags=re.M)
    return sorted(set(n for n in names if n.startswith("b200lap_") or n == "lapjv_seeded"))


@pytest.fixture(scope="module")
def built_lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200lap_build", os.path.join(ROOT, "gnn-accelerated-lap-warm-start-pipeline_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    path = mod.build()
    return ctypes.CDLL(path)


def test_header_declares_the_reference_symbol():
    syms = declared_symbols()
    assert "lapjv_seeded" in syms and len(syms) >= 25


def test_library_exports_every_declared_symbol(built_lib):
    for name in declared_symbols():
        assert hasattr(built_lib, name), f"libb200lap.so does not export {name}"


def test_ctypes_table_matches_header():
    from b200lap import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_a_device(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from b200lap import _lib
    lib = _lib.bind(built_lib)
    assert lib.b200lap_device_count() == 0
    C = np.array([[4.0, 1.0], [2.0, 0.5]])
    x = np.full(2, -1, np.int64)
    y = np.full(2, -1, np.int64)
    z = np.zeros(2)
    rc = lib.lapjv_seeded(C.ctypes.data, 2, 2, x.ctypes.data, y.ctypes.data, z.ctypes.data, z.ctypes.data, 1e-12)
    assert rc == _lib.ERR_CUDA and (x == -1).all() and (y == -1).all()
    # guards are the reference's own (lapjv_seeded.cpp:25-27) and come before any device work
    assert lib.lapjv_seeded(C.ctypes.data, 0, 0, x.ctypes.data, y.ctypes.data, z.ctypes.data, z.ctypes.data, 1e-12) == -2
    assert lib.lapjv_seeded(C.ctypes.data, 2, 1, x.ctypes.data, y.ctypes.data, z.ctypes.data, z.ctypes.data, 1e-12) == -4
    import lap
    with pytest.raises(RuntimeError):
        lap.lapjv_seeded(C, z, z)
    from b200lap import B200LapError, Context
    with pytest.raises(B200LapError):
        Context(0)


def test_python_binding_argument_checks():
    """Same rejections as the Cython signature of LAP/lap/_seeded_jv.pyx:14-25 (before any device work)."""
    import lap
    C = np.zeros((3, 3))
    with pytest.raises(ValueError):
        lap.lapjv_seeded(C, np.zeros(2), np.zeros(3))
    with pytest.raises(ValueError):
        lap.lapjv_seeded(C.astype(np.float32), np.zeros(3), np.zeros(3))
    with pytest.raises(ValueError):
        lap.lapjv_seeded(np.zeros((3, 3, 1)), np.zeros(3), np.zeros(3))
    with pytest.raises(ValueError):
        lap.lapjv_seeded(np.asfortranarray(np.arange(9.0).reshape(3, 3)), np.zeros(3), np.zeros(3))
    with pytest.raises(ValueError):
        lap.lapjv(np.zeros((2, 3)))
    with pytest.raises(ValueError):
        lap.lapjv(np.zeros(3))
