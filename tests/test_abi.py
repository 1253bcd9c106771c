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


def test_host_narrowing_of_the_upload(built_lib):
    """csrc/host_narrow.cpp marshals the host buffers of b200lap_pipeline_batch_submit (binary64 -> binary32 with an
    exactness check; the reference hands over binary64 matrices, scripts/gnn_benchmark.py:226).  Host-only code: checked
    here against numpy's own narrowing, including the cases that must be refused."""
    narrow = getattr(built_lib, "_ZN12b200lap_host6narrowEPKdPfmi")
    narrow.restype = ctypes.c_bool
    narrow.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rng = np.random.default_rng(0)
    for count in (1, 7, 8, 9, 1000, (1 << 20) + 3):
        a = rng.uniform(0, 1, count).astype(np.float32).astype(np.float64)
        for threads in (1, 3, 8):
            out = np.full(count, -1.0, np.float32)
            assert narrow(a.ctypes.data, out.ctypes.data, count, threads)
            assert np.array_equal(out, a.astype(np.float32)), (count, threads)
        out = np.empty(count, np.float32)
        b = a.copy(); b[count // 2] = 0.1            # not binary32-representable
        assert not narrow(b.ctypes.data, out.ctypes.data, count, 4)
        c = a.copy(); c[-1] = np.nan                 # NaN never survives the round-trip comparison
        assert not narrow(c.ctypes.data, out.ctypes.data, count, 4)
        d = a.copy(); d[0] = 1e300                   # overflows binary32
        assert not narrow(d.ctypes.data, out.ctypes.data, count, 2)
    t, p = ctypes.c_int(-1), ctypes.c_int(-1)
    built_lib.b200lap_host_narrow_config.restype = ctypes.c_longlong
    nbytes = built_lib.b200lap_host_narrow_config(64, 2048, ctypes.byref(t), ctypes.byref(p))
    assert t.value >= 0 and 0 <= p.value <= 100
    head = (64 * p.value + 50) // 100 if t.value > 0 else 0
    assert nbytes == 2048 * 2048 * (head * 4 + (64 - head) * 8)
