"""Checks and times b200lap_host::narrow (csrc/host_narrow.cpp) through the test build of the library: python tools/host_narrow_check.py"""
import ctypes, numpy as np, time, os
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "emul", "libb200lap_emul.so"))
f = getattr(lib, "_ZN12b200lap_host6narrowEPKdPfmi")
f.restype = ctypes.c_bool; f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
rng = np.random.default_rng(0)
for n in (1, 7, 8, 1000, 1 << 20, (1 << 24) + 3):
    a = rng.uniform(0, 1, n).astype(np.float32).astype(np.float64)
    out = np.empty(n, np.float32)
    for t in (1, 3, 8):
        assert f(a.ctypes.data, out.ctypes.data, n, t) and np.array_equal(out, a.astype(np.float32)), (n, t)
    b = a.copy(); b[n // 2] = 0.1        # not representable
    assert not f(b.ctypes.data, out.ctypes.data, n, 4), n
    c = a.copy(); c[-1] = np.nan
    assert not f(c.ctypes.data, out.ctypes.data, n, 4), n
a = rng.uniform(0, 1, 1 << 27).astype(np.float32).astype(np.float64); out = np.empty(1 << 27, np.float32)
for t in (1, 4, 8, 16, 24, 32):
    t0 = time.perf_counter(); f(a.ctypes.data, out.ctypes.data, a.size, t); dt = time.perf_counter() - t0
    print(t, "threads:", round(a.nbytes / dt / 1e9, 1), "GB/s read")
print("ok")
