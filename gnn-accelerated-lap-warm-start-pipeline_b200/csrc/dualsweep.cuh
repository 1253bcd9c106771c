// dualsweep.cuh -- the dual-potential sweeps of solvers/advanced_dual.py as kernels.
//
// Reference (NumPy statements, all binary64):
//   project_feasible  solvers/advanced_dual.py:14-36   u = min(u, min_j(C - v)); v = min(v, min_i(C - u)); until
//                                                      min(C - u - v) >= -tol or max_rounds
//   reduce_costs      solvers/advanced_dual.py:39-53   C' = (C - u) - v, shifted by -min(C') when that is negative
//   check_dual_feasible  :56-63                        min((C - u) - v) >= -tol
// Row minima min_j(c_ij - v_j) and the feasibility predicate ((c - u_i) - v_j < -tol) come from the solver's own
// front-end sweep (frontend.cuh, same expressions); the kernels here add the column sweep with binary64 row
// potentials, the element-wise clamps, the minimum reduced cost and the reduced-cost matrix.  Every quantity is a
// minimum or an element-wise expression evaluated exactly as NumPy evaluates it (left to right, no FMA), so results
// are bit-identical to the reference's.
#pragma once
#include "common.cuh"
#include "colsweep.cuh"

namespace b200lap {

// v_cap_j = min_i (c_ij - u_i), u binary64 (the min-trick kernel with binary64 row potentials)
template <typename CT, int VEC>
__global__ void __launch_bounds__(kColThreads) k_col_min_reduced_partial(
    const CT* __restrict__ C, long long inst_stride, int ld, int n, int rows_per_strip,
    const double* __restrict__ u /* [B][n] */, double* __restrict__ pval /* [B][S][n] */)
{
    const int b = blockIdx.z, strip = blockIdx.y, S = gridDim.y;
    const int c0 = (blockIdx.x * kColThreads + threadIdx.x) * VEC;
    if (c0 >= n) return;
    const CT* base = C + (size_t)b * inst_stride;
    const double* ub = u + (size_t)b * n;
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
            ui[q] = __ldg(ub + r + q);
        }
#pragma unroll
        for (int q = 0; q < kColUnroll; ++q)
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const double red = (double)t[q][e] - ui[q];
                best[e] = red < best[e] ? red : best[e];
            }
    }
    for (; r < r1; ++r) {
        CT t[VEC];
        VecLoad<CT, VEC>::ld(base + (size_t)r * ld + c0, t);
        const double ui = __ldg(ub + r);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const double red = (double)t[e] - ui;
            best[e] = red < best[e] ? red : best[e];
        }
    }
    const size_t o = ((size_t)b * S + strip) * n + c0;
#pragma unroll
    for (int e = 0; e < VEC; ++e)
        if (c0 + e < n) pval[o + e] = best[e];
}

// dst = min(dst, cap) element-wise (np.minimum)
__global__ void k_clamp_min(double* __restrict__ dst, const double* __restrict__ cap, long long count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) { const double a = dst[i], b = cap[i]; dst[i] = b < a ? b : a; }
}

// order-preserving image of a binary64 value for atomicMin on unsigned 64-bit words (and back)
__device__ __forceinline__ double ord2d(unsigned long long k) {
    const long long b = (long long)(k ^ (((long long)k < 0) ? 0x8000000000000000ull : 0xffffffffffffffffull));
    return __longlong_as_double(b);
}

// minimum reduced cost min_ij ((c_ij - u_i) - v_j) per instance -> min_ord[b] (ordered image, initialised to ~0ull),
// and optionally the reduced-cost matrix itself
template <typename CT, bool WRITE>
__global__ void __launch_bounds__(256) k_reduced_costs(
    const CT* __restrict__ C, long long inst_stride, int ld, int n, const double* __restrict__ u, const double* __restrict__ v,
    double* __restrict__ out /* [B][n][n] when WRITE */, unsigned long long* __restrict__ min_ord /* [B] */)
{
    __shared__ BlockRed s_red;
    const int b = blockIdx.y;
    const CT* base = C + (size_t)b * inst_stride;
    const double* ub = u + (size_t)b * n;
    const double* vb = v + (size_t)b * n;
    double m = INFINITY;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const double ui = ub[i];
        const CT* crow = base + (size_t)i * ld;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const double red = ((double)crow[j] - ui) - vb[j];
            if (WRITE) out[((size_t)b * n + i) * n + j] = red;
            m = red < m ? red : m;
        }
    }
    m = block_min_d(s_red, 0, m);
    if (threadIdx.x == 0 && m < INFINITY) atomicMin(min_ord + b, d2ord(m));
}

// out -= shift (np: Cprime - m)
__global__ void k_shift(double* __restrict__ out, long long count, double shift)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) out[i] = out[i] - shift;
}


// ---- oracle duals: the difference-constraint relaxation of solvers/dual_computation.py:13-47 --------------------
// The reference builds, for every matched pair (row r_p, column p), the edges p -> j of weight w = C[r_p, j] - C[r_p, p]
// and relaxes them Bellman-Ford style from v = 0 (`if v[b] > v[a] + w: v[b] = v[a] + w`) until nothing changes.  The
// operator is monotone, so any relaxation order reaches the same (greatest) fixed point; evaluating every edge as the
// reference does -- fl(v[a] + fl(c_ij - c_ip)) in binary64 -- makes that fixed point bit-identical.  Here one ROUND
// relaxes all n^2 edges at once (Jacobi): a column-minimum sweep of C with two per-row scalars.
__global__ void k_bf_gather(const void* __restrict__ C, int is_f64, long long inst_stride, int ld, int n, const int* __restrict__ x /* [B][n] row -> column */,
                            const double* __restrict__ v /* [B][n] */, double* __restrict__ vx, double* __restrict__ cx)
{
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = x[(size_t)b * n + i];
    const size_t o = (size_t)b * inst_stride + (size_t)i * ld + p;
    vx[(size_t)b * n + i] = v[(size_t)b * n + p];
    cx[(size_t)b * n + i] = is_f64 ? reinterpret_cast<const double*>(C)[o] : (double)reinterpret_cast<const float*>(C)[o];
}

template <typename CT, int VEC>
__global__ void __launch_bounds__(kColThreads) k_bf_relax_partial(
    const CT* __restrict__ C, long long inst_stride, int ld, int n, int rows_per_strip,
    const double* __restrict__ vx /* [B][n] v[x_i] */, const double* __restrict__ cx /* [B][n] c_{i, x_i} */, double* __restrict__ pval /* [B][S][n] */)
{
    const int b = blockIdx.z, strip = blockIdx.y, S = gridDim.y;
    const int c0 = (blockIdx.x * kColThreads + threadIdx.x) * VEC;
    if (c0 >= n) return;
    const CT* base = C + (size_t)b * inst_stride;
    const int r0 = strip * rows_per_strip;
    const int r1 = min(n, r0 + rows_per_strip);
    double best[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) best[e] = INFINITY;
    for (int r = r0; r < r1; ++r) {
        CT t[VEC];
        VecLoad<CT, VEC>::ld(base + (size_t)r * ld + c0, t);
        const double va = __ldg(vx + (size_t)b * n + r), ca = __ldg(cx + (size_t)b * n + r);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const double cand = va + ((double)t[e] - ca);
            best[e] = cand < best[e] ? cand : best[e];
        }
    }
    const size_t o = ((size_t)b * S + strip) * n + c0;
#pragma unroll
    for (int e = 0; e < VEC; ++e)
        if (c0 + e < n) pval[o + e] = best[e];
}

// v = min(v, cand); *changed |= something moved
__global__ void k_bf_update(double* __restrict__ v, const double* __restrict__ cand, long long count, int* __restrict__ changed)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) { const double a = v[i], c = cand[i]; if (a > c) { v[i] = c; *changed = 1; } }
}

}  // namespace b200lap
