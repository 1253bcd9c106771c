// frontend.cuh -- the warm-start front end of the seeded solve as ONE speculative sweep of C.
//
// Reference: LAP/_lapjv_cpp/lapjv_seeded.cpp:38-48 (Gauss-Seidel projection), :9-17,51
// (feasibility test), :66-73 (row tightening), :76-93 (greedy, consumes the tight lists built
// here), :105-113 (tight-edge count); solvers/advanced_dual.py:14-63 states the same sweeps in
// NumPy.
//
// The projection is a row-major chain, but it is the identity whenever no (i,j) violates
// u_i + v_j - c_ij <= eps under the SEED potentials (no trigger can fire before the first one).
// So one row-resident sweep computes, per row and with v = v_seed:
//     any violation            (u_seed_i + v_j) - c_ij  > eps        -> flags.any_viol
//     any infeasibility        (c_ij - u_seed_i) - v_j < -eps        -> flags.infeasible
//     u_i = min_j (c_ij - v_j)                                       -> u_tight
//     tight columns  |(c_ij - u_i) - v_j| <= tight_eps               -> count + first kTightCap, ascending
// and the solver kernel redoes the front end sequentially only when any_viol is set.
//
// Layout: a CTA owns `rows_per_cta` consecutive rows; every thread keeps its EPT columns of the
// current row AND of v in registers (C is read from HBM exactly once, v once per CTA), so the
// only on-chip traffic besides the reductions is the tight-column list.
#pragma once
#include "common.cuh"

namespace b200lap {

constexpr int kTightCap = 8;

struct FrontFlags {
    int any_viol;
    int infeasible;
    unsigned long long total_tight;
};

template <typename CT, int VEC> struct RowLoad;
template <> struct RowLoad<float, 4> {
    static __device__ __forceinline__ void ld(const float* p, float* o) {
        float4 t = __ldcs(reinterpret_cast<const float4*>(p));
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    }
};
template <> struct RowLoad<double, 2> {
    static __device__ __forceinline__ void ld(const double* p, double* o) {
        double2 t = __ldcs(reinterpret_cast<const double2*>(p));
        o[0] = t.x; o[1] = t.y;
    }
};
template <typename CT> struct RowLoad<CT, 1> {
    static __device__ __forceinline__ void ld(const CT* p, CT* o) { o[0] = __ldcs(p); }
};

// column owned by (thread, register slot e) in the VEC-interleaved row layout
template <int VEC> __device__ __forceinline__ int owned_col(int e, int T, int tid) { return ((e / VEC) * T + tid) * VEC + (e % VEC); }

// sort the first m (<= kTightCap) entries ascending (one thread, tiny)
__device__ __forceinline__ void sort_small(int* a, int m) {
    for (int p = 1; p < m; ++p) {
        int key = a[p], q = p - 1;
        while (q >= 0 && a[q] > key) { a[q + 1] = a[q]; --q; }
        a[q + 1] = key;
    }
}

// One barrier per row: the tight list of row r is double-buffered and finalised (sorted, written
// out) by thread 0 AFTER the min-reduction barrier of row r+1, when every thread has finished
// row r's tight pass.  VSM keeps v in shared memory instead of registers (long rows: 16 binary64
// registers per thread would spill under the 64-register cap of a 1024-thread CTA).
template <typename CT, int VEC, int EPT, bool VSM, int MAXT>
__global__ void __launch_bounds__(MAXT) k_front_end(
    const CT* __restrict__ C, long long inst_stride, int ld, int n, int rows_per_cta,
    const double* __restrict__ u_seed, const double* __restrict__ v_seed /* [B][n] */, double eps, double tight_eps,
    double* __restrict__ u_tight /* [B][n] */, int* __restrict__ tight_cols /* [B][n][kTightCap] */,
    int* __restrict__ tight_cnt /* [B][n] */, FrontFlags* __restrict__ flags /* [B] */)
{
    B200LAP_DYN_SMEM(dyn);
    __shared__ BlockRed s_red;
    __shared__ int s_cnt[2];
    __shared__ int s_list[2][kTightCap];
    const int b = blockIdx.y, T = blockDim.x, tid = threadIdx.x;
    const CT* base = C + (size_t)b * inst_stride;
    const double* vs = v_seed + (size_t)b * n;
    const double* us = u_seed + (size_t)b * n;
    double* sv = reinterpret_cast<double*>(dyn);
    if (tid < 2) s_cnt[tid] = 0;
    double vv[VSM ? 1 : EPT];
    if constexpr (VSM) {
        if constexpr (VEC == 4) {
            // split layout: a thread's 4 columns are two 16-byte loads, each conflict-free across the warp
            double2* sa = reinterpret_cast<double2*>(sv);
            double2* sb = sa + n / 4;
            for (int gi = tid; gi < n / 4; gi += T) {
                sa[gi] = make_double2(vs[4 * gi], vs[4 * gi + 1]);
                sb[gi] = make_double2(vs[4 * gi + 2], vs[4 * gi + 3]);
            }
        } else {
            for (int j = tid; j < n; j += T) sv[j] = vs[j];
        }
    } else {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int col = owned_col<VEC>(e, T, tid);
            vv[e] = col < n ? vs[col] : 0.0;
        }
    }
    __syncthreads();
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(n, r0 + rows_per_cta);
    int viol = 0, infeas = 0, par = 0, rpar = 0;
    unsigned long long total = 0;
    double ut_prev = 0.0;
    auto finalize = [&](int r, int p, double ut) {     // thread 0 only
        const int c = s_cnt[p];
        s_cnt[p] = 0;
        const int m2 = c < kTightCap ? c : kTightCap;
        sort_small(s_list[p], m2);
        int* out = tight_cols + ((size_t)b * n + r) * kTightCap;
        for (int q = 0; q < m2; ++q) out[q] = s_list[p][q];
        tight_cnt[(size_t)b * n + r] = c;
        u_tight[(size_t)b * n + r] = ut;
        total += (unsigned long long)c;
    };
    auto v_of = [&](int e, int col) -> double {
        if constexpr (!VSM) {
            return vv[e];
        } else if constexpr (VEC == 4) {
            const double2* sa = reinterpret_cast<const double2*>(sv);
            const double2 t = (e & 2) ? sa[n / 4 + (col >> 2)] : sa[col >> 2];   // the compiler merges the 4 uses of a group
            return (e & 1) ? t.y : t.x;
        } else {
            return sv[col];
        }
    };
    for (int r = r0; r < r1; ++r) {
        const CT* crow = base + (size_t)r * ld;
        CT cv[EPT];
#pragma unroll
        for (int g = 0; g < EPT / VEC; ++g) {
            const int col = owned_col<VEC>(g * VEC, T, tid);
            if (col < n) RowLoad<CT, VEC>::ld(crow + col, &cv[g * VEC]);
            else
#pragma unroll
                for (int q = 0; q < VEC; ++q) cv[g * VEC + q] = (CT)0;
        }
        const double ui = us[r];
        double m = INFINITY;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int col = owned_col<VEC>(e, T, tid);
            if (col < n) {
                const double c = (double)cv[e];
                const double vj = v_of(e, col);
                viol |= ((ui + vj) - c > eps);
                infeas |= ((c - ui) - vj < -eps);
                const double red = c - vj;
                m = red < m ? red : m;
            }
        }
        rpar ^= 1;
        const double ut = block_min_d(s_red, rpar, m);
        // every thread has finished the tight pass of row r-1: its list is complete
        if (tid == 0 && r > r0) finalize(r - 1, par ^ 1, ut_prev);
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int col = owned_col<VEC>(e, T, tid);
            if (col < n) {
                const double vj = v_of(e, col);
                const double rr = ((double)cv[e] - ut) - vj;
                if (fabs(rr) <= tight_eps) {
                    const int slot = atomicAdd(&s_cnt[par], 1);
                    if (slot < kTightCap) s_list[par][slot] = col;
                }
            }
        }
        ut_prev = ut;
        par ^= 1;
    }
    __syncthreads();
    if (tid == 0 && r1 > r0) finalize(r1 - 1, par ^ 1, ut_prev);
    if (viol) atomicOr(&flags[b].any_viol, 1);
    if (infeas) atomicOr(&flags[b].infeasible, 1);
    if (tid == 0 && total) atomicAdd(&flags[b].total_tight, total);
}

// ---- block-cooperative generic versions (used by the solver kernel when projection fires) -----
// One row of the Gauss-Seidel projection (lapjv_seeded.cpp:38-48): walk the columns in increasing order and, at every
// column that violates under the CURRENT u_i, apply the half/half update.  v lives in `v` (shared or global), u_i is
// returned.  Returns the number of updates.
//   * the whole CTA tests the row once under the incoming u_i (the common case: nothing fires, one reduction);
//   * when something fires, warp 0 alone walks the row from the first violation on, 32 columns at a time: a ballot
//     finds the next violating column, the owning lane's excess is broadcast, u_i and v_j move, the lanes behind it
//     re-test under the new u_i.  u_i only decreases, so columns that passed stay passed, and a trigger costs a dozen
//     warp instructions instead of a CTA-wide reduction with two barriers (seeds that violate broadly fire up to n^2
//     times: ~2 s per n = 2048 instance before, ~40 ms now).
template <typename CT>
__device__ int project_row(const CT* __restrict__ crow, int n, double* v, double& ui_io, double eps, Red& R)
{
    const int T = blockDim.x, tid = threadIdx.x;
    double ui = ui_io;
    int first = 0x7fffffff;
    for (int j = tid; j < n; j += T) {
        if ((ui + v[j]) - (double)crow[j] > eps) { first = j; break; }
    }
    const int jt = red_min_i(R, first);
    if (jt == 0x7fffffff) return 0;
    int fired = 0;
    if (warp_id() == 0) {
        const int lane = lane_id();
        for (int base = jt & ~31; base < n; base += 32) {
            const int j = base + lane;
            bool in = j < n && j >= jt;
            double vj = in ? v[j] : 0.0;
            const double cj = in ? (double)crow[j] : 0.0;
            while (true) {
                const double over = (ui + vj) - cj;
                const unsigned m = __ballot_sync(kFull, in && over > eps);
                if (!m) break;
                const int l = __ffs((int)m) - 1;
                const double half = __shfl_sync(kFull, over, l) / 2.0;
                ui -= half;
                if (lane == l) { vj -= half; v[j] = vj; }
                in = in && lane > l;
                ++fired;
            }
        }
    }
    // hand u_i and the count to everybody (also orders warp 0's writes of v before any later read)
    ui = red_min_d(R, warp_id() == 0 ? ui : INFINITY);
    fired = red_sum_i(R, tid == 0 ? fired : 0);
    ui_io = ui;
    return fired;
}

// Feasibility test + tightening + tight list of one row, generic column loop.
template <typename CT>
__device__ void front_row_generic(const CT* __restrict__ crow, int n, const double* v, double u_proj, double eps,
                                  double tight_eps, Red& R, int* s_cnt, int* s_list, int* tl_out, double* u_out,
                                  int* cnt_out, int* infeas_io)
{
    const int T = blockDim.x, tid = threadIdx.x;
    double m = INFINITY;
    int bad = 0;
    for (int j = tid; j < n; j += T) {
        const double c = (double)crow[j];
        bad |= ((c - u_proj) - v[j] < -eps);
        const double red = c - v[j];
        m = red < m ? red : m;
    }
    int badsum = 0;
    const double ut = red_min_d_sum_i(R, m, bad, &badsum);
    for (int j = tid; j < n; j += T) {
        const double rr = ((double)crow[j] - ut) - v[j];
        if (fabs(rr) <= tight_eps) {
            const int slot = atomicAdd(s_cnt, 1);
            if (slot < kTightCap) s_list[slot] = j;
        }
    }
    __syncthreads();
    const int c = *s_cnt;
    if (tid == 0) {
        const int m2 = c < kTightCap ? c : kTightCap;
        sort_small(s_list, m2);
        for (int q = 0; q < m2; ++q) tl_out[q] = s_list[q];
    }
    __syncthreads();
    if (tid == 0) *s_cnt = 0;
    *u_out = ut;
    *cnt_out = c;
    if (badsum) *infeas_io = 1;
}

}  // namespace b200lap
