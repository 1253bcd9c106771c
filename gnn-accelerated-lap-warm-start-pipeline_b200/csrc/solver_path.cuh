// solver_path.cuh -- one shortest augmenting path (find_path_dense) with the Dijkstra state in REGISTERS and
// BATCHED relax steps over the SCAN queue.
//
// Reference: LAP/_lapjv_cpp/lapjv.cpp:153-171 (_find_dense), :178-213 (_scan_dense), :221-282 (find_path_dense).
//
// Measurements behind the design (B200, tools/micro/lat.cu, tools/batch_breakdown.py; profiles/r02_*):
//   * a random 8 KB matrix row costs ~950 cycles from DRAM with 64 instances in flight, ~500 from L2;
//   * the per-column arithmetic is pipe bound on ONE SM: F2F.F64.F32 16 lanes/clk, DSETP 32 lanes/clk, DADD 64 lanes/clk
//     -> ~330 cycles per scanned row of n = 2048;
//   * 60 % of the relax steps of the bench batch start with >= 5 columns already waiting in SCAN (the tight-edge
//     search is wide: structural ties after every dual update), and a CTA barrier + shared-memory round trip is ~100.
// So:
//   * a thread owns MAXC fixed columns for the whole path -- groups of VEC = 16 / sizeof(CT) consecutive columns, one
//     128-bit load per group and row -- with distance d and potential v in registers and two bit masks, `todo` (still
//     in the TODO zone of cols[]) and `ready` (READY at the last level collect: its potential moves at the end);
//   * a relax step takes up to KMAX columns from SCAN at once: their rows are fetched together (one memory latency),
//     relaxed one after the other in registers (exactly the reference's sequential semantics per column: a column hit
//     by an earlier scan of the batch has left TODO for the later ones), and all hits (cand == level) are published
//     with (scan index, column, row y[j], potential) through a 3-slot rotating list;
//   * after ONE barrier warp 0 replays the hits scan by scan, each scan's hits in ascending position order (positions
//     are read after the swaps of the earlier scans) -- the first unmatched column ends the path exactly where the
//     reference's early return does, later scans of the batch then never happened (their register updates are dead
//     state: d/pred of TODO columns are re-initialised by the next path, READY columns are never touched);
//   * every column that enters SCAN gets its queue entry (row, potential) written by position and its matrix row
//     PREFETCHED INTO L2 (cp.async.bulk.prefetch.L2) -- the queue is deep, so by the time the row is scanned it is an
//     L2 hit instead of a DRAM access;
//   * a level collect transposes d from column order (registers) to POSITION order (the shared array `d`), flags the
//     prefix-minimum records, and warp 0 replays them from (position, column, distance) tuples.
// All state arrays are addressed as true shared memory here (the path is only taken when everything fits).
#pragma once
#include "common.cuh"

namespace b200lap {

// Measurement build only: step timings of two observer threads (thread 0, thread 33) in trace words
// 20 + 10 * observer + stamp.  A stamp is taken when `dep` is available (the predicate makes the clock read wait).
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

// L2 prefetch of one matrix row (bytes a multiple of 16, 16-byte aligned address)
__device__ __forceinline__ void prefetch_row_l2(const void* p, unsigned bytes) {
#ifndef B200LAP_EMUL
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
    (void)p; (void)bytes;
#endif
}

// the state arrays as TRUE shared-memory pointers (same packing order as k_solve's placement with a full mask)
struct SmemState {
    unsigned int* bitmap;
    double *d, *v;
    int *pos, *y, *cols, *pred, *x, *free_rows, *srow;
};
__device__ __forceinline__ SmemState smem_state(unsigned char* dyn, int n) {
    SmemState m;
    size_t off = 0;
    auto take = [&](int arr) { unsigned char* p = dyn + off; off += state_array_bytes(arr, n); return p; };
    m.bitmap = (unsigned int*)take(ST_BITMAP);
    m.d = (double*)take(ST_D);
    m.pos = (int*)take(ST_POS);
    m.v = (double*)take(ST_V);
    m.y = (int*)take(ST_Y);
    m.cols = (int*)take(ST_COLS);
    m.pred = (int*)take(ST_PRED);
    m.x = (int*)take(ST_X);
    m.free_rows = (int*)take(ST_FREE);
    m.srow = (int*)take(ST_SROW);
    return m;
}

template <typename CT> struct PathCtx {
    const CT* C;
    int ld, n;
    unsigned row_bytes;
    SmemState m;
    SolverShared* sh;
};

// a column enters SCAN at position k: queue entry + L2 prefetch of the row it will be scanned through
template <typename CT>
__device__ __forceinline__ void scan_enqueue(const PathCtx<CT>& P, int k, int yj, double vj) {
    P.m.srow[k] = yj;
    P.m.d[k] = vj;            // `d` holds distances by position only during a collect; positions in SCAN carry the potential
    if (yj >= 0) prefetch_row_l2(P.C + (size_t)yj * P.ld, P.row_bytes);
}

// ---- serial replay of a level collect, d in POSITION order (warp 0) -------------------------------------------
// The flagged positions (prefix-minimum records and ties, ascending) are turned into (position, column, distance)
// tuples by the whole warp -- a record's position is not touched by the swaps of earlier records, so column and
// distance can be fetched up front -- and lane 0 then replays the reference's swaps over the tuples.
template <typename CT>
__device__ __forceinline__ void replay_collect_pos(const PathCtx<CT>& P, int lo, int wlo, int whi)
{
    const int lane = lane_id();
    const long long t0 = sm_clock();
    SolverShared* sh = P.sh;
    const SmemState& m = P.m;
    int hi = lo, total = 0;
    double level = INFINITY;
    for (int w0 = wlo; w0 <= whi; w0 += 32) {
        unsigned int bits = 0;
        if (w0 + lane <= whi) { bits = m.bitmap[w0 + lane]; if (bits) m.bitmap[w0 + lane] = 0u; }
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
            const int cntc = min(kRecCap, round_total - c0);
            for (int r = lane; r < cntc; r += 32) {
                const int k = sh->rec[r].k;
                sh->rec[r].j = m.cols[k];
                sh->rec[r].d = m.d[k];
            }
            __syncwarp();
            if (lane == 0) {
                RecTuple t = sh->rec[0];
                for (int r = 0; r < cntc; ++r) {
                    const RecTuple nx = sh->rec[r + 1 < cntc ? r + 1 : r];
                    if (t.d < level) { hi = lo; level = t.d; }
                    const int c2 = m.cols[hi];
                    m.cols[t.k] = c2; m.pos[c2] = t.k;
                    m.cols[hi] = t.j; m.pos[t.j] = hi;
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
    // unmatched column among the collected level: the LAST one in position order wins (lapjv.cpp:250-255);
    // every collected column gets its SCAN-queue entry
    int best = -1;
    for (int k = lo + lane; k < hi; k += 32) {
        const int j = m.cols[k];
        const int yj = m.y[j];
        if (yj < 0) best = k;
        scan_enqueue(P, k, yj, m.v[j]);
    }
    best = warp_max_i(best);
    if (lane == 0) {
        sh->hi = hi;
        sh->final_j = best >= 0 ? m.cols[best] : -1;
        sh->level = level;
        B200LAP_PROF(sh->tr[TR_RECORDS] += total; sh->tr[TR_CYC_COLLECT_REPLAY] += sm_clock() - t0);
    }
}

// ---- replay of the hits of a batched relax step from the published per-scan lists (warp 0) ----------------------
// Scan by scan; within a scan ascending position (read after the swaps of the earlier scans); the first unmatched
// column ends everything (lapjv.cpp:199-203).  A scan with a single hit -- the common case -- needs no ordering at
// all: every lane reads the same entry and lane 0 performs the swap.  Results go to sh->res[parity].
template <typename CT>
__device__ __forceinline__ void replay_hits_list(const PathCtx<CT>& P, int hi_in, int K, int slot, int parity)
{
    const int lane = lane_id();
    SolverShared* sh = P.sh;
    const SmemState& m = P.m;
    int hi = hi_in, fin = -1, done = K;
    for (int s = 0; s < K; ++s) {
        const int c = sh->nh[slot][s];
        if (c == 0) continue;
        __syncwarp();
        if (c == 1) {
            const HitEntry e = sh->hl[slot][s][0];
            const int j = e.sj;
            if (e.yj < 0) { fin = j; done = s + 1; break; }
            if (lane == 0) {
                const int k = m.pos[j];
                const int c2 = m.cols[hi];
                m.cols[k] = c2; m.pos[c2] = k;
                m.cols[hi] = j; m.pos[j] = hi;
                scan_enqueue(P, hi, e.yj, e.v);
            }
            ++hi;
            continue;
        }
        HitEntry e;
        e.sj = 0; e.yj = 0; e.v = 0.0;
        if (lane < c) e = sh->hl[slot][s][lane];
        int my_k = lane < c ? m.pos[e.sj] : 0x7fffffff;
        for (int q = 0; q < c; ++q) {
            __syncwarp();
            const int kmin = __reduce_min_sync(kFull, my_k);
            const int src = __ffs((int)__ballot_sync(kFull, my_k == kmin)) - 1;
            const int j = __shfl_sync(kFull, e.sj, src);
            const int yj = __shfl_sync(kFull, e.yj, src);
            if (yj < 0) { fin = j; break; }
            if (lane == src) {
                const int c2 = m.cols[hi];
                m.cols[kmin] = c2; m.pos[c2] = kmin;
                m.cols[hi] = j; m.pos[j] = hi;
                scan_enqueue(P, hi, yj, e.v);
                my_k = 0x7fffffff;
            }
            ++hi;
        }
        if (fin >= 0) { done = s + 1; break; }
    }
    __syncwarp();
    if (lane == 0) { sh->res[parity].hi = hi; sh->res[parity].final_j = fin; sh->res[parity].done = done; }
}

// ---- replay of ONE scan's hits from the position bitmap (overflow path: more than kHitCap hits in a batch) ------
template <typename CT>
__device__ __forceinline__ void replay_hits_bitmap(const PathCtx<CT>& P, int hi_in, int wlo, int whi)
{
    const int lane = lane_id();
    SolverShared* sh = P.sh;
    const SmemState& m = P.m;
    int hi = hi_in, fin = -1;
    for (int w0 = wlo; w0 <= whi; w0 += 32) {
        unsigned int bits = 0;
        if (w0 + lane <= whi) { bits = m.bitmap[w0 + lane]; m.bitmap[w0 + lane] = 0u; }
        unsigned int nz = __ballot_sync(kFull, bits != 0u);
        while (nz) {
            const int l = __ffs((int)nz) - 1;
            nz &= nz - 1;
            unsigned int word = __shfl_sync(kFull, bits, l);
            if (lane == 0 && fin < 0) {
                while (word) {
                    const int bpos = __ffs((int)word) - 1;
                    word &= word - 1;
                    const int k = (w0 + l) * 32 + bpos;
                    const int j = m.cols[k];
                    const int yj = m.y[j];
                    if (yj < 0) { fin = j; break; }
                    const int c2 = m.cols[hi];
                    m.cols[k] = c2; m.pos[c2] = k;
                    m.cols[hi] = j; m.pos[j] = hi;
                    scan_enqueue(P, hi, yj, m.v[j]);
                    ++hi;
                }
            }
        }
    }
    if (lane == 0) { sh->hi = hi; sh->final_j = fin; }
}

template <typename CT> struct VecOf;
template <> struct VecOf<float> {
    static constexpr int V = 4;
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

// Requires: n % VEC == 0, rows 16-byte aligned, blockDim.x * MAXC >= n, single CTA, ALL state arrays in shared memory.
template <int MAXC, typename CT>
__device__ int shortest_path_reg(SolverCtx<CT>& S, int start_i)
{
    typedef VecOf<CT> VT;
    constexpr int V = VT::V;
    constexpr int G = MAXC / V;
    constexpr int KMAX = (4 * kMaxScans) / MAXC > kMaxScans ? kMaxScans : ((4 * kMaxScans) / MAXC < 1 ? 1 : (4 * kMaxScans) / MAXC);    // scans per batch: KMAX * MAXC row entries in registers
    static_assert(MAXC % V == 0 && MAXC <= 16, "register-resident path: MAXC must be a multiple of the vector width");
    B200LAP_DYN_SMEM(dyn);
    const int n = S.n, T = blockDim.x, tid = threadIdx.x;
    SolverShared* sh = S.sh;
    PathCtx<CT> P;
    P.C = S.C; P.ld = S.ld; P.n = n; P.row_bytes = (unsigned)(n * sizeof(CT)); P.m = smem_state(dyn, n); P.sh = sh;
    const SmemState& m = P.m;
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
                VT::ld_d(m.v + base, &vq[g * V]);
                VT::st_idx(m.cols + base, base);
                VT::st_idx(m.pos + base, base);
                VT::st_val(m.pred + base, start_i);
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
    int n_collect = 0, n_relax = 0;      // trace counters (uniform), flushed by thread 0 at the end of the path
    // hits of the previous batch that warp 0 has not replayed yet (uniform): it replays them in the shadow of the NEXT
    // batch's row fetch whenever SCAN still holds columns that were queued before
    bool pend = false;
    int pend_K = 0, pend_sp = 0, batch_no = 0;
#if defined(B200LAP_SOLVER_PROFILE) && !defined(B200LAP_EMUL)
    const int obs = tid == 0 ? 0 : (tid == 33 ? 1 : -1);
    long long t_last = 0;
#endif
    for (;;) {
        if (lo == hi && pend) {
            // SCAN ran dry: the pending hits decide how it goes on
            if (warp_id() == 0) replay_hits_list(P, hi, pend_K, pend_sp, batch_no & 1);
            __syncthreads();
            const HitResult r = sh->res[batch_no & 1];
            ++batch_no;
            pend = false;
            hi = r.hi;
            n_relax += r.done;
            if (r.final_j >= 0) { final_j = r.final_j; break; }
        }
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
                    VT::ld_idx(m.pos + base, pk);
#pragma unroll
                    for (int q = 0; q < V; ++q)
                        if ((todo >> (g * V + q)) & 1u) m.d[pk[q]] = dq[g * V + q];
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
            for (int q = 0; q < MAXC; ++q) dch[q] = (q < chunk && k0 + q < k1) ? m.d[k0 + q] : INFINITY;
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
                    atomicOr(&m.bitmap[k >> 5], 1u << (k & 31));
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
#pragma unroll
                    for (int s = 0; s < kMaxScans; ++s) sh->nh[old_slot][s] = 0;
                }
                if (whi < 0) { wlo = lo >> 5; whi = wlo; }
                replay_collect_pos(P, lo, wlo, whi);
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
        }
        // ================= one BATCHED relax step (_scan_dense over up to KMAX SCAN columns) =================
        const long long tr0 = sm_clock();
#if defined(B200LAP_SOLVER_PROFILE) && !defined(B200LAP_EMUL)
        t_last = tr0;
#endif
        const int K = min(min(KMAX, S.kcap), hi - lo);
        const int sp = S.step % 3;
        int ej[KMAX], er[KMAX];
        double ev[KMAX];
#pragma unroll
        for (int s = 0; s < KMAX; ++s) {
            const int k = lo + (s < K ? s : 0);
            ej[s] = m.cols[k];
            er[s] = m.srow[k];
            ev[s] = m.d[k];
        }
        B200LAP_PROF(if (tid == 0) { const int dep_ = hi - lo; sh->tr[40 + (dep_ <= 1 ? 0 : dep_ == 2 ? 1 : dep_ <= 4 ? 2 : 3)] += 1; });
        B200LAP_STAMP(0, er[0] >= 0 && ev[0] == ev[0]);
        CT cr[KMAX][MAXC];
        CT cjs[KMAX];
#pragma unroll
        for (int s = 0; s < KMAX; ++s) {
            if (s < K) {
                const CT* crow = S.C + (size_t)er[s] * S.ld;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int base = (g * T + tid) * V;
                    if (base < n) VT::ld(crow + base, &cr[s][g * V]);
                }
                cjs[s] = __ldg(crow + ej[s]);
            }
        }
        B200LAP_STAMP(1, true);
        // the previous batch's hits: swaps, queue entries and row prefetches, while this batch's rows are in flight
        if (pend && warp_id() == 0) replay_hits_list(P, hi, pend_K, pend_sp, batch_no & 1);
        unsigned hitm = 0;
#pragma unroll
        for (int s = 0; s < KMAX; ++s) {
            if (s < K) {
                const double slack = ((double)cjs[s] - ev[s]) - level;
                if (s == 0) { B200LAP_STAMP(2, slack == slack); }
#pragma unroll
                for (int e = 0; e < MAXC; ++e) {
                    const int col = ((e / V) * T + tid) * V + (e % V);
                    const double cand = ((double)cr[s][e] - vq[e]) - slack;
                    if (((todo >> e) & 1u) && cand < dq[e]) {
                        dq[e] = cand;
                        m.pred[col] = er[s];
                        if (cand == level) { hitm |= 1u << (s * MAXC + e); todo &= ~(1u << e); }
                    }
                }
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
        // publish the hits per scan: column, its row and its potential
        for (unsigned hb = hitm; hb; hb &= hb - 1) {
            const int b = __ffs((int)hb) - 1;
            const int s = b / MAXC, e = b % MAXC;
            const int col = ((e / V) * T + tid) * V + (e % V);
            const int slot = atomicAdd(&sh->nh[sp][s], 1);
            if (slot < kHitCap) {
                HitEntry he;
                he.sj = col; he.yj = m.y[col]; he.v = m.v[col];
                sh->hl[sp][s][slot] = he;
            }
        }
        B200LAP_STAMP(4, true);
        __syncthreads();
        B200LAP_STAMP(5, true);
        S.step++;
        int cnt = 0, cmax = 0;
#pragma unroll
        for (int s = 0; s < KMAX; ++s) { const int c = sh->nh[sp][s]; cnt += c; cmax = max(cmax, c); }
        if (tid == 0) {
            const int old_slot = (sp + 2) % 3;
            S.minw[old_slot] = 0x7fffffff; S.maxw[old_slot] = -1; S.nhit[old_slot] = 0;
#pragma unroll
            for (int s = 0; s < kMaxScans; ++s) sh->nh[old_slot][s] = 0;
        }
        B200LAP_STAMP(6, cnt >= 0);
        B200LAP_PROF(if (tid == 0) { sh->tr[44 + (cnt == 0 ? 0 : cnt == 1 ? 1 : 2)] += 1; if (cnt > 1) sh->tr[47] += cnt; });
        if (pend) {
            // what the replay of the previous batch found; if it ended the path, this batch never happened
            const HitResult r = sh->res[batch_no & 1];
            ++batch_no;
            pend = false;
            hi = r.hi;
            n_relax += r.done;
            if (r.final_j >= 0) { final_j = r.final_j; break; }
        }
        lo += K;
        if (cnt == 0) {
            n_relax += K;
        } else if (cmax <= kHitCap) {
            pend = true; pend_K = K; pend_sp = sp;
            if (!S.pipe) {
                if (warp_id() == 0) replay_hits_list(P, hi, pend_K, pend_sp, batch_no & 1);
                __syncthreads();
                const HitResult r = sh->res[batch_no & 1];
                ++batch_no;
                pend = false;
                hi = r.hi;
                n_relax += r.done;
                if (r.final_j >= 0) { final_j = r.final_j; break; }
            }
        } else {
            // more hits in one scan than its list holds (tie-heavy instances): one scan at a time through the position bitmap
            int hi_cur = hi, done = K;
#pragma unroll 1
            for (int s = 0; s < K; ++s) {
                int wmin_i = 0x7fffffff, wmax_i = -1;
                for (unsigned hb = (hitm >> (s * MAXC)) & ((1u << MAXC) - 1u); hb; hb &= hb - 1) {
                    const int e = __ffs((int)hb) - 1;
                    const int col = ((e / V) * T + tid) * V + (e % V);
                    const int k = m.pos[col];
                    atomicOr(&m.bitmap[k >> 5], 1u << (k & 31));
                    wmin_i = min(wmin_i, k >> 5);
                    wmax_i = max(wmax_i, k >> 5);
                }
                if (wmax_i >= 0) { atomicMin(&S.minw[sp], wmin_i); atomicMax(&S.maxw[sp], wmax_i); }
                __syncthreads();
                if (warp_id() == 0) {
                    const int wlo = S.minw[sp], whi = S.maxw[sp];
                    if (whi >= 0) replay_hits_bitmap(P, hi_cur, wlo, whi);
                    else if (lane_id() == 0) { sh->hi = hi_cur; sh->final_j = -1; }
                    __syncwarp();
                    if (lane_id() == 0) { S.minw[sp] = 0x7fffffff; S.maxw[sp] = -1; }
                }
                __syncthreads();
                hi_cur = sh->hi;
                final_j = sh->final_j;
                if (final_j >= 0) { done = s + 1; break; }
            }
            hi = hi_cur;
            n_relax += done;
            if (final_j >= 0) break;
        }
        B200LAP_STAMP(7, hi >= 0);
        B200LAP_PROF(if (tid == 0) sh->tr[TR_CYC_RELAX] += sm_clock() - tr0);
    }
    // ---- dual update of the READY columns (lapjv.cpp:270-276): level == d[cols[n_ready]]
#pragma unroll
    for (int e = 0; e < MAXC; ++e) {
        if ((ready >> e) & 1u) {
            const int col = ((e / V) * T + tid) * V + (e % V);
            m.v[col] = vq[e] + (dq[e] - level);
        }
    }
    if (tid == 0) { sh->tr[TR_COLLECT] += n_collect; sh->tr[TR_RELAX] += n_relax; }
    return final_j;
}

}  // namespace b200lap
