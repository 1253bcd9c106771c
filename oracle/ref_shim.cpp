// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// The reference's lapjv_internal/_ccrrt_dense/_carr_dense/_ca_dense have C++
// linkage (LAP/_lapjv_cpp/lapjv.h:59-69), so ctypes cannot bind them directly.
// This shim is compiled *together with* the untouched reference sources
// (see oracle/Makefile: they are compiled where they lie under /root/reference)
// and re-exports them with C linkage and flat row-major matrices.
#include <cstdlib>
#include <vector>

extern int lapjv_internal(const unsigned int n, double *cost[], int *x, int *y);

extern "C" int ref_lapjv_internal(const double *C, int n, int *x, int *y)
{
    if (n <= 0) return -2;
    std::vector<double *> rows((size_t)n);
    for (int i = 0; i < n; ++i) rows[(size_t)i] = const_cast<double *>(C) + (size_t)i * (size_t)n;
    return lapjv_internal((unsigned int)n, rows.data(), x, y);
}
