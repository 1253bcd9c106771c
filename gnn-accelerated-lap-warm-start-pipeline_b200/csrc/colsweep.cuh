// colsweep.cuh -- column-wise reductions over a row-major cost matrix (one read of C).
//
//   col_argmin  : per column the minimum raw cost and the FIRST row attaining it.  Feeds the
//                 is_col_best feature (gnn/features.py:218-219) and the column-reduction step of
//                 the cold solve (LAP/_lapjv_cpp/lapjv.cpp:21-32: row-major sweep, strict '<').
//   min_trick   : v_j = min_i (C_ij - u_i) in binary64 on widened operands
//                 (scripts/gnn_benchmark.py:262; min is order independent => bit exact).
//
// Layout: grid = (column tiles, row strips, instances).  A thread owns VEC adjacent columns
// (one 128-bit load per row when VEC*sizeof(CT)==16) and walks the rows of its strip with
// UNROLL independent loads in flight; strips write (value,row) partials that a small
// finalize kernel folds in ascending strip order (keeps the first-row rule).
#pragma once
#include "common.cuh"

namespace b200lap {

constexpr int kColThreads = 256;
constexpr int kColUnroll = 8;

template <typename CT, int VEC> struct VecLoad;
template <> struct VecLoad<float, 4> {
    static __device__ __forceinline__ void ld(const float* p, float (&o)[4]) {
        float4 t = __ldcs(reinterpret_cast<const float4*>(p));
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    }
};
template <> struct VecLoad<double, 2> {
    static __device__ __forceinline__ void ld(const double* p, double (&o)[2]) {
        double2 t = __ldcs(reinterpret_cast<const double2*>(p));
        o[0] = t.x; o[1] = t.y;
    }
};
template <typename CT> struct VecLoad<CT, 1> {
    static __device__ __forceinline__ void ld(const CT* p, CT (&o)[1]) { o[0] = __ldcs(p); }
};

// ---- col_argmin ---------------------------------------------------------------------------
template <typename CT, int VEC>
__global__ void __launch_bounds__(kColThreads) k_col_argmin_partial(
    const CT* __restrict__ C, long long inst_stride, int ld, int n, int rows_per_strip,
    CT* __restrict__ pval, int* __restrict__ prow /* [B][S][n] */)
{
    const int b = blockIdx.z, strip = blockIdx.y, S = gridDim.y;
    const int c0 = (blockIdx.x * kColThreads + threadIdx.x) * VEC;
    if (c0 >= n) return;
    const CT* base = C + (size_t)b * inst_stride;
    const int r0 = strip * rows_per_strip;
    const int r1 = min(n, r0 + rows_per_strip);
    CT best[VEC];
    int brow[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { best[e] = (CT)INFINITY; brow[e] = -1; }
    int r = r0;
    for (; r + kColUnroll <= r1; r += kColUnroll) {
        CT t[kColUnroll][VEC];
#pragma unroll
        for (int q = 0; q < kColUnroll; ++q) VecLoad<CT, VEC>::ld(base + (size_t)(r + q) * ld + c0, t[q]);
#pragma unroll
        for (int q = 0; q < kColUnroll; ++q)
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                if (t[q][e] < best[e]) { best[e] = t[q][e]; brow[e] = r + q; }
    }
    for (; r < r1; ++r) {
        CT t[VEC];
        VecLoad<CT, VEC>::ld(base + (size_t)r * ld + c0, t);
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            if (t[e] < best[e]) { best[e] = t[e]; brow[e] = r; }
    }
    const size_t o = ((size_t)b * S + strip) * n + c0;
#pragma unroll
    for (int e = 0; e < VEC; ++e)
        if (c0 + e < n) { pval[o + e] = best[e]; prow[o + e] = brow[e]; }
}

template <typename CT>
__global__ void k_col_argmin_final(const CT* __restrict__ pval, const int* __restrict__ prow, int S, int n,
                                   CT* __restrict__ colmin, int* __restrict__ colarg /* [B][n] */)
{
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    CT best = (CT)INFINITY;
    int brow = -1;
    int s = 0;
    for (; s + 8 <= S; s += 8) {       // 16 independent loads in flight, folded in strip order
        CT pv[8];
        int pr[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const size_t o = ((size_t)b * S + s + q) * n + j;
            pv[q] = pval[o];
            pr[q] = prow[o];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (pv[q] < best) { best = pv[q]; brow = pr[q]; }
    }
    for (; s < S; ++s) {
        const size_t o = ((size_t)b * S + s) * n + j;
        CT v = pval[o];
        if (v < best) { best = v; brow = prow[o]; }
    }
    colmin[(size_t)b * n + j] = best;
    colarg[(size_t)b * n + j] = brow;
}

// ---- min-trick ------------------------------------------------------------------------------
template <typename CT, int VEC>
__global__ void __launch_bounds__(kColThreads) k_min_trick_partial(
    const CT* __restrict__ C, long long inst_stride, int ld, int n, int rows_per_strip,
    const float* __restrict__ u /* [B][n] binary32 row potentials */, double* __restrict__ pval /* [B][S][n] */)
{
    const int b = blockIdx.z, strip = blockIdx.y, S = gridDim.y;
    const int c0 = (blockIdx.x * kColThreads + threadIdx.x) * VEC;
    if (c0 >= n) return;
    const CT* base = C + (size_t)b * inst_stride;
    const float* ub = u + (size_t)b * n;
    const int r0 = strip * rows_per_strip;
    const int r1 = min(n, r0 + rows_per_strip);
    double best[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) best[e] = INFINITY;
    int r = r0;
    for (; r + kColUnroll <= r1; r += kColUnroll) {
        CT t[kColUnroll][VEC];
        double ui[kColUnroll];
#pragma unroll
        for (int q = 0; q < kColUnroll; ++q) {
            VecLoad<CT, VEC>::ld(base + (size_t)(r + q) * ld + c0, t[q]);
            ui[q] = (double)__ldg(ub + r + q);
        }
#pragma unroll
        for (int q = 0; q < kColUnroll; ++q)
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                double red = (double)t[q][e] - ui[q];
                best[e] = red < best[e] ? red : best[e];
            }
    }
    for (; r < r1; ++r) {
        CT t[VEC];
        VecLoad<CT, VEC>::ld(base + (size_t)r * ld + c0, t);
        const double ui = (double)__ldg(ub + r);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            double red = (double)t[e] - ui;
            best[e] = red < best[e] ? red : best[e];
        }
    }
    const size_t o = ((size_t)b * S + strip) * n + c0;
#pragma unroll
    for (int e = 0; e < VEC; ++e)
        if (c0 + e < n) pval[o + e] = best[e];
}

__global__ void k_min_trick_final(const double* __restrict__ pval, int S, int n, double* __restrict__ v /* [B][n] */)
{
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double best = INFINITY;
#pragma unroll 8
    for (int s = 0; s < S; ++s) {
        double t = pval[((size_t)b * S + s) * n + j];
        best = t < best ? t : best;
    }
    v[(size_t)b * n + j] = best;
}

}  // namespace b200lap
