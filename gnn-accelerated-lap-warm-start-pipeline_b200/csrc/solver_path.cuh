// solver_path.cuh -- one shortest augmenting path (find_path_dense) with the Dijkstra state in REGISTERS.
//
// Reference: LAP/_lapjv_cpp/lapjv.cpp:153-171 (_find_dense), :178-213 (_scan_dense), :221-282 (find_path_dense).
//
// Layout.  A thread owns MAXC fixed columns for the whole path -- groups of VEC = 16 / sizeof(CT) consecutive
// columns, one 128-bit load per group and matrix row -- and keeps their distance d and potential v in registers,
// together with two bit masks: `todo` (the column is still in the TODO zone of cols[]) and `ready` (it was READY at
// the last level collect, i.e. its potential moves at the end of the path).  A relax step (_scan_dense) is then
//     1 row load per group  ->  cand = (c - v) - slack, compare with d, select   (no shared-memory read per column)
// and the only things that leave the thread are the predecessor of an improved column (a predicated store) and the
// rare hit (cand == level), published through a 3-slot rotating mailbox.  The common step -- exactly one hit -- is
// finished by EVERY thread redundantly from the mailbox (hit column, its row y[j], its potential), so the next
// row fetch is issued right after the step's single barrier; the swap of cols[]/pos[] that the reference performs
// for the hit is done by thread 0 one step late, in the shadow of that fetch (nothing reads cols/pos in between:
// hits are published by column, positions are only looked up after a barrier).  Steps with several hits, and the
// level collects, keep the position bitmap + serial replay of solver.cuh (bit-exact order of cols[]).
//
// A level collect transposes d from column order (registers) to POSITION order (the shared array S.d is used as
// d-by-position here), so that the prefix-minimum scan over positions [lo, n) reads consecutive words.
#pragma once
#include "common.cuh"

namespace b200lap {

// Measurement build only: step timings of two observer threads (thread 0 = the bookkeeping thread, thread 33 = an
// ordinary one) in trace words 20 + 10 * observer + stamp.  A stamp is taken when `dep` is available (the predicate
// makes the clock read wait for it).
#if defined(B200LAP_SOLVER_PROFILE) && !defined(B200LAP_EMUL)
#define B200LAP_STAMP(idx, dep)                                                         \
    do {                                                                                \
        if (obs >= 0 && (dep)) {                                                        \
            const long long t_ = clock64();                                             \
            sh->tr[20 + 10 * obs + (idx)] += t_ - t_last;                               \
            t_last = t_;                                                                \
        }                                                                               \
    } while (0)
#else
#define B200LAP_STAMP(idx, dep) do { } while (0)
#endif

// ---- serial replay of a level collect, d in POSITION order (warp 0) -------------------------------------------
// The flagged positions (prefix-minimum records and ties, ascending) are turned into (position, column, distance)
// tuples by the whole warp -- a record's position is not touched by the swaps of earlier records, so column and
// distance can be fetched up front -- and lane 0 then replays the reference's swaps over the tuples with nothing
// but one dependent shared-memory load per record on its chain (the next tuple is already in registers).
template <typename CT>
__device__ __forceinline__ void replay_collect_pos(SolverCtx<CT>& S, int lo, int wlo, int whi)
{
    const int lane = lane_id();
    const long long t0 = sm_clock();
    SolverShared* sh = S.sh;
    int hi = lo, total = 0;
    double level = INFINITY;
    for (int w0 = wlo; w0 <= whi; w0 += 32) {
        unsigned int bits = 0;
        if (w0 + lane <= whi) { bits = S.bitmap[w0 + lane]; if (bits) S.bitmap[w0 + lane] = 0u; }
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        const int round_total = __shfl_sync(kFull, incl, 31);
        for (int c0 = 0; c0 < round_total; c0 += kRecCap) {
            int idx = incl - cnt - c0;
            unsigned int w = bits;
            while (w) {
                const int bpos = __ffs((int)w) - 1;
                w &= w - 1;
                if (idx >= 0 && idx < kRecCap) sh->rec[idx].k = (w0 + lane) * 32 + bpos;
                ++idx;
            }
            __syncwarp();
            const int m = min(kRecCap, round_total - c0);
            for (int r = lane; r < m; r += 32) {
                const int k = sh->rec[r].k;
                sh->rec[r].j = S.cols[k];
                sh->rec[r].d = S.d[k];
            }
            __syncwarp();
            if (lane == 0) {
                RecTuple t = sh->rec[0];
                for (int r = 0; r < m; ++r) {
                    const RecTuple nx = sh->rec[r + 1 < m ? r + 1 : r];
                    if (t.d < level) { hi = lo; level = t.d; }
                    const int c2 = S.cols[hi];
                    S.cols[t.k] = c2; S.pos[c2] = t.k;
                    S.cols[hi] = t.j; S.pos[t.j] = hi;
                    ++hi;
                    t = nx;
                }
            }
            __syncwarp();
        }
        total += round_total;
    }
    hi = __shfl_sync(kFull, hi, 0);
    if (hi == lo) hi = lo + 1;   // only reachable with NaN distances; keep moving
    // unmatched column among the collected level: the LAST one in position order wins (lapjv.cpp:250-255)
    int best = -1;
    for (int k = lo + lane; k < hi; k += 32)
        if (S.y[S.cols[k]] < 0) best = k;
    best = warp_max_i(best);
    if (lane == 0) {
        sh->hi = hi;
        sh->final_j = best >= 0 ? S.cols[best] : -1;
        sh->level = level;
        B200LAP_PROF(sh->tr[TR_RECORDS] += total; sh->tr[TR_CYC_COLLECT_REPLAY] += sm_clock() - t0);
    }
}

template <typename CT> struct VecOf;
template <> struct VecOf<float> {
    static constexpr int V = 4;
    typedef float4 type;
    static __device__ __forceinline__ void ld(const float* p, float* o) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    }
    static __device__ __forceinline__ void st_idx(int* p, int a) { *reinterpret_cast<int4*>(p) = int4{a, a + 1, a + 2, a + 3}; }
    static __device__ __forceinline__ void st_val(int* p, int a) { *reinterpret_cast<int4*>(p) = int4{a, a, a, a}; }
    static __device__ __forceinline__ void ld_idx(const int* p, int* o) {
        const int4 t = *reinterpret_cast<const int4*>(p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    }
    static __device__ __forceinline__ void ld_d(const double* p, double* o) {
        const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
};
template <> struct VecOf<double> {
    static constexpr int V = 2;
    typedef double2 type;
    static __device__ __forceinline__ void ld(const double* p, double* o) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p));
        o[0] = t.x; o[1] = t.y;
    }
    static __device__ __forceinline__ void st_idx(int* p, int a) { *reinterpret_cast<int2*>(p) = int2{a, a + 1}; }
    static __device__ __forceinline__ void st_val(int* p, int a) { *reinterpret_cast<int2*>(p) = int2{a, a}; }
    static __device__ __forceinline__ void ld_idx(const int* p, int* o) {
        const int2 t = *reinterpret_cast<const int2*>(p);
        o[0] = t.x; o[1] = t.y;
    }
    static __device__ __forceinline__ void ld_d(const double* p, double* o) {
        const double2 a = *reinterpret_cast<const double2*>(p);
        o[0] = a.x; o[1] = a.y;
    }
};

// Requires: n % VEC == 0, rows 16-byte aligned, blockDim.x * MAXC >= n, single CTA (S.nc == 1), all state arrays
// addressable (shared or global -- only S.v, S.y are read and S.pred written per step, the rest at collects).
template <int MAXC, typename CT>
__device__ int shortest_path_reg(SolverCtx<CT>& S, int start_i)
{
    typedef VecOf<CT> VT;
    constexpr int V = VT::V;
    constexpr int G = MAXC / V;
    static_assert(MAXC % V == 0 && MAXC <= 32, "register-resident path: MAXC must be a multiple of the vector width");
    const int n = S.n, T = blockDim.x, tid = threadIdx.x;
    SolverShared* sh = S.sh;
    double dq[MAXC], vq[MAXC];
    unsigned todo = 0, ready = 0, valid = 0;
    // ---- start of the path: d = C[start] - v, identity permutation, every predecessor the root
    {
        const CT* row0 = S.C + (size_t)start_i * S.ld;
        CT c0[MAXC];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int base = (g * T + tid) * V;
            if (base < n) VT::ld(row0 + base, &c0[g * V]);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int base = (g * T + tid) * V;
            if (base < n) {
                VT::ld_d(S.v + base, &vq[g * V]);
                VT::st_idx(S.cols + base, base);
                VT::st_idx(S.pos + base, base);
                VT::st_val(S.pred + base, start_i);
                valid |= ((1u << V) - 1u) << (g * V);
            } else {
#pragma unroll
                for (int q = 0; q < V; ++q) vq[g * V + q] = 0.0;
            }
#pragma unroll
            for (int q = 0; q < V; ++q) dq[g * V + q] = base < n ? (double)c0[g * V + q] - vq[g * V + q] : INFINITY;
        }
        todo = valid;
    }
    int lo = 0, hi = 0, final_j = -1;
    double level = 0.0;
    int js = 0, irow = 0;
    double vjs = 0.0;
    bool have_entry = false;
    int pend_j = -1, pend_hi = 0;        // thread 0: the swap of the previous step's single hit, not yet applied
    int n_collect = 0, n_relax = 0;      // thread 0: trace counters, flushed at the end of the path
#if defined(B200LAP_SOLVER_PROFILE) && !defined(B200LAP_EMUL)
    const int obs = tid == 0 ? 0 : (tid == 33 ? 1 : -1);
    long long t_last = 0;
#endif
    for (;;) {
        if (lo == hi) {
            // ================= level collect (_find_dense) =================
            const long long tc0 = sm_clock();
            ready |= valid & ~todo;          // everything that left TODO so far sits below n_ready = lo
            // d in position order for the prefix-minimum scan
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int base = (g * T + tid) * V;
                if (base < n && ((todo >> (g * V)) & ((1u << V) - 1u))) {
                    int pk[V];
                    VT::ld_idx(S.pos + base, pk);
#pragma unroll
                    for (int q = 0; q < V; ++q)
                        if ((todo >> (g * V + q)) & 1u) S.d[pk[q]] = dq[g * V + q];
                }
            }
            __syncthreads();
            B200LAP_PROF(if (tid == 0) sh->tr[28] += sm_clock() - tc0);
            const long long tc1 = sm_clock();
            const int L = n - lo;
            const int chunk = (L + T - 1) / T;
            const int k0 = lo + tid * chunk;
            const int k1 = min(n, k0 + chunk);
            double dch[MAXC];
            double lm = INFINITY;
#pragma unroll
            for (int q = 0; q < MAXC; ++q) dch[q] = (q < chunk && k0 + q < k1) ? S.d[k0 + q] : INFINITY;
#pragma unroll
            for (int q = 0; q < MAXC; ++q) lm = dch[q] < lm ? dch[q] : lm;
            double incl = lm;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t = __shfl_up_sync(kFull, incl, o);
                if (lane_id() >= o) incl = t < incl ? t : incl;
            }
            double before = __shfl_up_sync(kFull, incl, 1);
            if (lane_id() == 0) before = INFINITY;
            const int p = S.R.flip();
            if (lane_id() == 31) S.R.r->d[p][warp_id()] = incl;
            __syncthreads();
            B200LAP_PROF(if (tid == 0) sh->tr[29] += sm_clock() - tc1);
            const long long tc2 = sm_clock();
            {
                const double t = lane_id() < warp_id() ? S.R.r->d[p][lane_id()] : INFINITY;
                const double wmin = warp_min_d(t);
                before = wmin < before ? wmin : before;
            }
            const int sp = S.step % 3;
            int wmin_i = 0x7fffffff, wmax_i = -1;
            double run = before;
#pragma unroll
            for (int q = 0; q < MAXC; ++q) {
                if (q < chunk && k0 + q < k1 && dch[q] <= run) {
                    const int k = k0 + q;
                    atomicOr(&S.bitmap[k >> 5], 1u << (k & 31));
                    wmin_i = min(wmin_i, k >> 5);
                    wmax_i = max(wmax_i, k >> 5);
                    run = dch[q];
                }
            }
            if (wmax_i >= 0) { atomicMin(&S.minw[sp], wmin_i); atomicMax(&S.maxw[sp], wmax_i); }
            __syncthreads();
            B200LAP_PROF(if (tid == 0) sh->tr[38] += sm_clock() - tc2);
            const long long tc3 = sm_clock();
            if (warp_id() == 0) {
                int wlo = S.minw[sp], whi = S.maxw[sp];
                if (lane_id() == 0) {
                    const int old_slot = (sp + 2) % 3;
                    S.minw[old_slot] = 0x7fffffff; S.maxw[old_slot] = -1; S.nhit[old_slot] = 0;
                }
                if (whi < 0) { wlo = lo >> 5; whi = wlo; }
                replay_collect_pos(S, lo, wlo, whi);
            }
            __syncthreads();
            B200LAP_PROF(if (tid == 0) sh->tr[39] += sm_clock() - tc3);
            S.step++;
            ++n_collect;
            hi = sh->hi;
            final_j = sh->final_j;
            level = sh->level;
            // the columns at the new level are SCAN now
#pragma unroll
            for (int e = 0; e < MAXC; ++e)
                if (((todo >> e) & 1u) && dq[e] == level) todo &= ~(1u << e);
            B200LAP_PROF(if (tid == 0) sh->tr[TR_CYC_COLLECT] += sm_clock() - tc0);
            if (final_j >= 0) break;
            have_entry = false;
        }
        // ================= one relax step (_scan_dense, one SCAN column) =================
        const long long tr0 = sm_clock();
#if defined(B200LAP_SOLVER_PROFILE) && !defined(B200LAP_EMUL)
        t_last = tr0;
#endif
        if (!have_entry) {
            js = S.cols[lo];
            irow = S.y[js];
            vjs = S.v[js];
        }
        const int sp = S.step % 3;
        const CT* crow = S.C + (size_t)irow * S.ld;
        B200LAP_STAMP(0, irow >= 0 && vjs == vjs);
        CT cr[MAXC];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int base = (g * T + tid) * V;
            if (base < n) VT::ld(crow + base, &cr[g * V]);
        }
        const CT c_js = __ldg(crow + js);
        if (tid == 0 && pend_j >= 0) {
            // the reference's swap for the previous step's hit: cols[k] = cols[hi]; cols[hi++] = j
            const int k = S.pos[pend_j];
            const int c2 = S.cols[pend_hi];
            S.cols[k] = c2; S.pos[c2] = k;
            S.cols[pend_hi] = pend_j; S.pos[pend_j] = pend_hi;
            pend_j = -1;
        }
        B200LAP_STAMP(1, pend_j < 0);
        const double slack = ((double)c_js - vjs) - level;
        // profile build: word 13 = cycles until the row arrived (the slack needs c_js), word 14 = cycles in the barrier
        B200LAP_PROF(if (tid == 0 && slack == slack) sh->tr[TR_CYC_ARR_SCAN] += sm_clock() - tr0);
        B200LAP_STAMP(2, slack == slack);
        unsigned hitm = 0;
#pragma unroll
        for (int e = 0; e < MAXC; ++e) {
            const int col = ((e / V) * T + tid) * V + (e % V);
            const double cand = ((double)cr[e] - vq[e]) - slack;
            if (((todo >> e) & 1u) && cand < dq[e]) {
                dq[e] = cand;
                S.pred[col] = irow;
                if (cand == level) hitm |= 1u << e;
            }
        }
#if defined(B200LAP_SOLVER_PROFILE) && !defined(B200LAP_EMUL)
        {
            double acc_ = 0.0;
#pragma unroll
            for (int e = 0; e < MAXC; ++e) acc_ += dq[e];
            B200LAP_STAMP(3, acc_ == acc_ || hitm);
        }
#endif
        if (hitm) {
            todo &= ~hitm;
#pragma unroll
            for (int e = 0; e < MAXC; ++e) {
                if ((hitm >> e) & 1u) {
                    const int col = ((e / V) * T + tid) * V + (e % V);
                    if (atomicAdd(&S.nhit[sp], 1) == 0) {
                        S.hit_j[sp] = col;
                        sh->hit_y[sp] = S.y[col];
                        sh->hit_v[sp] = vq[e];
                    }
                }
            }
        }
        const long long tb0 = sm_clock();
        B200LAP_STAMP(4, true);
        __syncthreads();
        B200LAP_STAMP(5, true);
        B200LAP_PROF(if (tid == 0) sh->tr[TR_CYC_ARR_SERIAL] += sm_clock() - tb0);
        S.step++;
        const int nh = S.nhit[sp];
        B200LAP_STAMP(6, nh >= 0);
        if (tid == 0) {
            const int old_slot = (sp + 2) % 3;
            S.minw[old_slot] = 0x7fffffff; S.maxw[old_slot] = -1; S.nhit[old_slot] = 0;
            ++n_relax;
        }
        ++lo;
        have_entry = false;
        if (nh == 1) {
            const int j = S.hit_j[sp], yj = sh->hit_y[sp];
            B200LAP_PROF(if (tid == 0) sh->tr[TR_RELAX_HITS] += 1);
            if (yj < 0) { final_j = j; break; }
            if (tid == 0) { pend_j = j; pend_hi = hi; }
            if (lo == hi) { js = j; irow = yj; vjs = sh->hit_v[sp]; have_entry = true; }
            ++hi;
        } else if (nh > 1) {
            // several hits: order them by position (bitmap) and replay serially, as the reference scans k ascending
            int wmin_i = 0x7fffffff, wmax_i = -1;
#pragma unroll
            for (int e = 0; e < MAXC; ++e) {
                if ((hitm >> e) & 1u) {
                    const int col = ((e / V) * T + tid) * V + (e % V);
                    const int k = S.pos[col];
                    atomicOr(&S.bitmap[k >> 5], 1u << (k & 31));
                    wmin_i = min(wmin_i, k >> 5);
                    wmax_i = max(wmax_i, k >> 5);
                }
            }
            if (wmax_i >= 0) { atomicMin(&S.minw[sp], wmin_i); atomicMax(&S.maxw[sp], wmax_i); }
            __syncthreads();
            if (warp_id() == 0) replay_relax(S, hi, S.minw[sp], S.maxw[sp]);
            __syncthreads();
            hi = sh->hi;
            final_j = sh->final_j;
            if (final_j >= 0) break;
        }
        B200LAP_STAMP(7, hi >= 0 && irow >= 0);
        B200LAP_PROF(if (tid == 0) sh->tr[TR_CYC_RELAX] += sm_clock() - tr0);
    }
    // ---- dual update of the READY columns (lapjv.cpp:270-276): level == d[cols[n_ready]]
#pragma unroll
    for (int e = 0; e < MAXC; ++e) {
        if ((ready >> e) & 1u) {
            const int col = ((e / V) * T + tid) * V + (e % V);
            S.v[col] = vq[e] + (dq[e] - level);
        }
    }
    if (tid == 0) { sh->tr[TR_COLLECT] += n_collect; sh->tr[TR_RELAX] += n_relax; }
    return final_j;
}

}  // namespace b200lap
