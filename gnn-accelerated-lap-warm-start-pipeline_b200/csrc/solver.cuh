// solver.cuh -- the seeded Jonker-Volgenant solve as ONE persistent CTA per instance.
//
// Reference: LAP/_lapjv_cpp/lapjv_seeded.cpp:19-173 (lapjv_seeded) and
// LAP/_lapjv_cpp/lapjv.cpp:8-346 (_ccrrt_dense, _carr_dense, _find_dense, _scan_dense,
// find_path_dense, _ca_dense, lapjv_internal).  Everything after the front-end sweep
// (frontend.cuh) and the column-argmin sweep (colsweep.cuh) runs inside k_solve: greedy
// first-fit over the tight lists, the 1.2 n fallback test, micro-ARR, Dijkstra augmentation,
// and -- for the fallback -- column reduction, reduction transfer, two ARR passes and
// augmentation.  No kernel launch per Dijkstra/ARR step: a step is a coalesced gather of one
// row of C, a warp-shuffle + shared-memory reduction, and a short serial replay.
//
// Bit-exactness: all arithmetic is binary64 on widened matrix entries, written with the
// reference's association, compiled with -fmad=false.  The history-dependent order of the
// cols[] permutation (every prefix-minimum record and every tie swaps, lapjv.cpp:153-171; hits
// are appended in position order and the first unmatched hit ends the scan, :178-213) is
// reproduced by flagging hits in parallel into a POSITION bitmap and replaying the flagged
// positions serially in ascending order: a swap only moves an already-visited, unflagged column
// to an already-visited position, so pending flags stay valid.
//
// State (per instance): v,d binary64[n]; pred,cols,pos,y,x,free_rows int32[n]; bitmap n/32
// words.  It lives in shared memory when 40 n + n/8 bytes fit (n <= ~5600), else in a
// global-memory workspace that stays L2-resident.
#pragma once
#include "common.cuh"
#include "frontend.cuh"

namespace b200lap {

constexpr int kTraceWords = 48;   // == B200LAP_TRACE_WORDS; words 20.. are fine-grained step timings of the measurement build
enum TraceSlot {
    TR_PROJ = 0, TR_TIGHT = 1, TR_GREEDY = 2, TR_FALLBACK = 3, TR_MICRO = 4, TR_FREE_CR = 5,
    TR_ARR = 6, TR_PATHS = 7, TR_COLLECT = 8, TR_RELAX = 9, TR_RC = 10,
    // SM-clock cycles spent (thread 0's view) in: relax steps, collect steps, ARR row scans + reductions, ARR serial updates, total
    TR_CYC_RELAX = 11, TR_CYC_COLLECT = 12, TR_CYC_ARR_SCAN = 13, TR_CYC_ARR_SERIAL = 14, TR_CYC_TOTAL = 15,
    // records replayed by collect steps, cycles of the serial replays (warp 0), hits replayed by relax steps
    TR_RECORDS = 16, TR_CYC_COLLECT_REPLAY = 17, TR_CYC_RELAX_REPLAY = 18, TR_RELAX_HITS = 19
};

struct RecTuple { int k, j; double d; };   // one record of a level collect: position, column, distance
constexpr int kRecCap = 128;
struct HitEntry { int sj, yj; double v; }; // one hit of a batched relax step: column, the row matched to it, its potential
struct HitResult { int hi, final_j, done, pad; };   // what the replay of a batch's hits found
constexpr int kHitCap = 8;                 // hits per scan the lists hold (more: position-bitmap path)
#ifndef B200LAP_KMAX
#define B200LAP_KMAX 4
#endif
constexpr int kMaxScans = B200LAP_KMAX;    // scans per batched relax step

struct SolverShared {
    BlockRed red;
    BlockRed2 red2;
    int s_cnt;
    int s_list[kTightCap];
    int hi, final_j, next_row, aux;
    int minw[3], maxw[3];   // hit-word ranges, rotated over 3 steps (reset one step after use)
    int nhit[3], hit_k[3], hit_j[3];   // relax steps: number of hits and the first one published (same rotation)
    int hit_y[3];                      // register-resident path: row matched to the published hit column, and its potential
    double hit_v[3];
    double level;                      // register-resident path: the level a collect step ended with
    RecTuple rec[kRecCap];             // register-resident path: the records of a collect step, in position order
    HitEntry hl[3][kMaxScans][kHitCap];   // register-resident path: hits of a batched relax step per scan (3-slot rotation as nhit)
    int nh[3][kMaxScans];              // register-resident path: hits per scan
    HitResult res[2];                  // register-resident path: replay results, double-buffered by batch parity
    int box_op, box_row, box_js, box_hi, box_sp;   // cluster mode: the master's command mailbox, read by the workers through DSMEM
    int box_k, box_rows[kMaxScans], box_jss[kMaxScans];   // cluster mode: the scans of one batched relax step (row, scanned column)
    int bnhit[3][kMaxScans], bhit_j[3][kMaxScans];        // cluster mode: hits per scan and the first hit column published (3-slot rotation)
    int bminw[3][kMaxScans], bmaxw[3][kMaxScans];         // cluster mode: word range of the scan's hits in its COLUMN bitmap
    unsigned int cursor, deferred;
    int hitk[64];           // positions of the flagged records of a collect step, ascending
    long long tr[kTraceWords];
};

template <typename CT> struct SolveArgs {
    const CT* C;
    long long inst_stride;
    int ld, n;
    const double* u_seed;   // [B][n]
    const double* v_seed;   // [B][n]
    double eps;
    int mode;               // 0 = seeded (lapjv_seeded), 1 = cold (lapjv_internal only)
    const double* u_tight;  // [B][n]        (front-end sweep)
    int* tight_cols;        // [B][n][kTightCap]
    int* tight_cnt;         // [B][n]
    const FrontFlags* flags;// [B]
    const CT* colmin;       // [B][n]        (column-argmin sweep)
    const int* colarg;      // [B][n]
    unsigned char* gws;     // global state workspace, `gws_stride` bytes per instance (or null)
    long long gws_stride;
    int smem_mask;          // bit a set: state array a lives in shared memory (order of kStateArrays), else in gws
    int* x;                 // [B][n] out
    int* y;                 // [B][n] out
    int* rc;                // [B] out
    long long* trace;       // [B][kTraceWords] out (nullable)
    double* v_out;          // [B][n] final column potentials (nullable)
    int cluster;            // CTAs per instance (thread-block cluster size), 1 = single CTA
    int regpath;            // 1: augmentation keeps d/v in registers (solver_path.cuh); needs vector-aligned rows
    int kcap;               // register path: scans per batched relax step (0 = as many as the registers hold)
    int pipe;               // register path: replay a batch's hits behind the next batch's row fetch
};

// State arrays in placement priority order (hottest first): a relax step reads d, pos, v of every
// column; the replay and the path flip touch y, cols, pred, x sparsely.  Whatever fits goes to shared
// memory, the rest to the (L2-resident) global workspace -- n = 8192 keeps d, pos, v, y on chip,
// n = 16384 keeps d and pos.
enum StateArray { ST_D = 0, ST_POS, ST_V, ST_Y, ST_COLS, ST_PRED, ST_X, ST_FREE, ST_BITMAP, ST_SROW, ST_COUNT };
__host__ __device__ inline size_t state_array_bytes(int a, int n) {
    const size_t n8 = ((size_t)n + 7) & ~(size_t)7;
    if (a == ST_D || a == ST_V) return n8 * 8;
    if (a == ST_BITMAP) return ((((size_t)n + 31) / 32 + 4) * 4 + 15) & ~(size_t)15;
    return n8 * 4;
}
// greedy placement under a shared-memory budget -> (mask, shared bytes)
// `allowed`: arrays that may live in shared memory at all.  In cluster mode only the arrays the master alone
// touches (y, cols, x, free_rows) qualify, plus v, of which the workers read a global mirror refreshed at the
// start of every path (v only changes between paths); d, pos, pred and the bitmap are shared with the workers.
constexpr int kClusterSmemArrays = (1 << ST_V) | (1 << ST_Y) | (1 << ST_COLS) | (1 << ST_X) | (1 << ST_FREE);
__host__ inline int solver_place_state(int n, size_t budget, size_t* smem_bytes, int allowed = (1 << ST_COUNT) - 1) {
    int mask = 0;
    size_t used = 0;
    const int order_all[ST_COUNT] = {ST_BITMAP, ST_D, ST_POS, ST_V, ST_Y, ST_COLS, ST_PRED, ST_X, ST_FREE, ST_SROW};
    const int order_cluster[ST_COUNT] = {ST_Y, ST_COLS, ST_V, ST_X, ST_FREE, ST_BITMAP, ST_D, ST_POS, ST_PRED, ST_SROW};
    const int* order = allowed == (1 << ST_COUNT) - 1 ? order_all : order_cluster;
    for (int q = 0; q < ST_COUNT; ++q) {
        const size_t b = state_array_bytes(order[q], n);
        if ((allowed >> order[q] & 1) && used + b <= budget) { used += b; mask |= 1 << order[q]; }
    }
    *smem_bytes = used;
    return mask;
}

__host__ __device__ inline size_t solver_state_bytes(int n) {
    const size_t n8 = ((size_t)n + 7) & ~(size_t)7;
    return n8 * 8 * 2 + n8 * 4 * 7 + (((size_t)n + 31) / 32 + 4) * 4 + 16;
}

// ---- cluster mode (large instances): the relax step of the Dijkstra search is spread over the CTAs of a
// thread-block cluster.  CTA 0 (the master) runs the solver as usual on state kept in the L2-resident global
// workspace; at each relax step it posts (row, scanned column, hi, slot) in a mailbox IN ITS SHARED MEMORY, which the
// workers read through distributed shared memory (ld.shared::cluster); the cluster barrier
// releases the workers, every CTA relaxes its own slice of columns, a second cluster barrier publishes the
// d/pred updates and the hit flags (hit slots: atom/st.shared::cluster on the master's shared memory), and the master
// replays the hits from its own shared memory.  Workers sit in worker_loop.
enum { BOX_RELAX = 1, BOX_EXIT = 2 };
// distributed-shared-memory accessors (32-bit shared::cluster addresses; a CTA may address its own window this way too)
#ifndef B200LAP_EMUL
__device__ __forceinline__ unsigned dsm_map(const void* local_smem, unsigned cta_rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ int dsm_ld(unsigned addr) {
    int v;
    asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void dsm_st(unsigned addr, int v) { asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ int dsm_add(unsigned addr, int v) {
    int o;
    asm volatile("atom.shared::cluster.add.s32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v) : "memory");
    return o;
}
__device__ __forceinline__ void dsm_min(unsigned addr, int v) { int o; asm volatile("atom.shared::cluster.min.s32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void dsm_max(unsigned addr, int v) { int o; asm volatile("atom.shared::cluster.max.s32 %0, [%1], %2;" : "=r"(o) : "r"(addr), "r"(v) : "memory"); }
#else
__device__ __forceinline__ unsigned dsm_map(const void*, unsigned) { return 0; }
__device__ __forceinline__ int dsm_ld(unsigned) { return 0; }
__device__ __forceinline__ void dsm_st(unsigned, int) {}
__device__ __forceinline__ int dsm_add(unsigned, int) { return 0; }
__device__ __forceinline__ void dsm_min(unsigned, int) {}
__device__ __forceinline__ void dsm_max(unsigned, int) {}
#endif
__device__ __forceinline__ void cluster_sync_all() {
#ifndef B200LAP_EMUL
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}

template <typename CT> struct SolverCtx {
    const CT* C;
    int ld, n;
    double *v, *d;
    double* vg;   // what the worker CTAs read: v itself, or its global mirror when v lives in the master's shared memory
    int *pred, *cols, *pos, *y, *x, *free_rows;
    int* srow;    // register-resident path: row matched to the column at each SCAN-queue position
    unsigned int* bitmap;
    SolverShared* sh;
    Red R;
    int step;   // relax/collect step counter (selects the minw/maxw slot)
    int *minw, *maxw, *nhit, *hit_k, *hit_j;   // slot arrays: in SolverShared, or in the cluster mailbox (cluster mode)
    int nc, rank;                              // thread-block cluster size and this CTA's rank (1, 0 without a cluster)
    unsigned msh;                              // cluster mode: shared::cluster address of the MASTER's SolverShared
    unsigned int* cbm;                         // cluster mode: per-scan COLUMN bitmaps of a batched relax step (global workspace)
    int cbw;                                   // words per scan in cbm
    int regpath;                               // augmentation with register-resident d/v (solver_path.cuh)
    int kcap, pipe;
};

// Cycle-level phase counters (trace words 11..19) are compiled in only with -DB200LAP_SOLVER_PROFILE
// (B200LAP_PROFILE=1 python build.py; tools/solver_breakdown.py): their shared-memory read-modify-writes sit on
// the serial chain of every step.  Without the macro those trace words stay zero.
#if defined(B200LAP_SOLVER_PROFILE) && !defined(B200LAP_EMUL)
#define B200LAP_PROF(stmt) do { stmt; } while (0)
__device__ __forceinline__ long long sm_clock() { return clock64(); }
#else
#define B200LAP_PROF(stmt) do { } while (0)
__device__ __forceinline__ long long sm_clock() { return 0; }
#endif

// ---- one thread-strided pass over a matrix row ----------------------------------------------------
// With MAXC > 0 (blockDim.x * MAXC >= n) the thread's entries are loaded into registers FIRST, as
// independent coalesced loads, and only then handed to `body` -- a row costs one memory latency, not
// one per entry (the bodies branch on shared state, which otherwise serialises the loads).
template <int MAXC, typename CT, typename F>
__device__ __forceinline__ void row_scan(const CT* __restrict__ crow, int n, F&& body)
{
    const int T = blockDim.x, tid = threadIdx.x;
    if constexpr (MAXC > 0) {
        CT c[MAXC];
#pragma unroll
        for (int q = 0; q < MAXC; ++q) {
            const int j = tid + q * T;
            c[q] = j < n ? __ldg(crow + j) : (CT)0;
        }
#pragma unroll
        for (int q = 0; q < MAXC; ++q) {
            const int j = tid + q * T;
            if (j < n) body(j, (double)c[q]);
        }
    } else {
        for (int j = tid; j < n; j += T) body(j, (double)crow[j]);
    }
}

// ---- serial replay of flagged positions (warp 0) ------------------------------------------------
// mode 0: level collect (_find_dense): every flagged position is a prefix-minimum record or tie.
// mode 1: relax (_scan_dense): every flagged position reached the level; the first unmatched
//         one ends the path search.
template <typename CT>
__device__ __forceinline__ void replay_collect(SolverCtx<CT>& S, int lo, int wlo, int whi)
{
    // executed by warp 0 only
    const int lane = lane_id();
    const long long t0 = sm_clock();
    int hi = lo;
    double level = INFINITY;
    // 1) enumerate the flagged positions (ascending) into hitk[]: all lanes work, 32 words per round
    int total = 0;
    bool overflow = false;
    for (int w0 = wlo; w0 <= whi; w0 += 32) {
        const unsigned int bits = (w0 + lane <= whi) ? S.bitmap[w0 + lane] : 0u;
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        const int round_total = __shfl_sync(kFull, incl, 31);
        if (total + round_total > 64) { overflow = true; break; }
        int slot = total + incl - cnt;
        unsigned int w = bits;
        while (w) {
            const int bpos = __ffs((int)w) - 1;
            w &= w - 1;
            S.sh->hitk[slot++] = (w0 + lane) * 32 + bpos;
        }
        total += round_total;
    }
    __syncwarp();
    if (!overflow) {
        // 2) fetch (position, column, distance) of every record IN PARALLEL -- none of them is touched by the
        //    swaps of earlier records -- and leave only the cols[hi] read on lane 0's serial chain
        for (int w = wlo + lane; w <= whi; w += 32) S.bitmap[w] = 0u;
        for (int base = 0; base < total; base += 32) {
            const int cnt = min(32, total - base);
            int my_k = 0, my_j = 0;
            double my_d = INFINITY;
            if (lane < cnt) { my_k = S.sh->hitk[base + lane]; my_j = S.cols[my_k]; my_d = S.d[my_j]; }
            for (int h = 0; h < cnt; ++h) {
                const int k = __shfl_sync(kFull, my_k, h);
                const int j = __shfl_sync(kFull, my_j, h);
                const double dj = __shfl_sync(kFull, my_d, h);
                if (lane == 0) {
                    if (dj < level) { hi = lo; level = dj; }
                    const int c2 = S.cols[hi];
                    S.cols[k] = c2; S.pos[c2] = k;
                    S.cols[hi] = j; S.pos[j] = hi;
                    ++hi;
                }
            }
        }
    } else {
        // tie-heavy level (more than 64 records): plain serial walk over the bitmap
        total = 0;
        for (int w0 = wlo; w0 <= whi; w0 += 32) {
            unsigned int bits = 0;
            if (w0 + lane <= whi) { bits = S.bitmap[w0 + lane]; S.bitmap[w0 + lane] = 0u; }
            unsigned int nz = __ballot_sync(kFull, bits != 0u);
            while (nz) {
                const int l = __ffs((int)nz) - 1;
                nz &= nz - 1;
                unsigned int word = __shfl_sync(kFull, bits, l);
                if (lane == 0) {
                    while (word) {
                        const int bpos = __ffs((int)word) - 1;
                        word &= word - 1;
                        const int k = (w0 + l) * 32 + bpos;
                        const int j = S.cols[k];
                        const double dj = S.d[j];
                        if (dj < level) { hi = lo; level = dj; }
                        const int c2 = S.cols[hi];
                        S.cols[k] = c2; S.pos[c2] = k;
                        S.cols[hi] = j; S.pos[j] = hi;
                        ++hi;
                        ++total;
                    }
                }
            }
        }
        total = __shfl_sync(kFull, total, 0);
    }
    hi = __shfl_sync(kFull, hi, 0);
    if (hi == lo) hi = lo + 1;   // only reachable with NaN distances; keep moving
    __syncwarp();
    // unmatched column among the collected level: the LAST one in position order wins (lapjv.cpp:250-255)
    int best = -1;
    for (int k = lo + lane; k < hi; k += 32)
        if (S.y[S.cols[k]] < 0) best = k;
    best = warp_max_i(best);
    if (lane == 0) {
        S.sh->hi = hi;
        S.sh->final_j = best >= 0 ? S.cols[best] : -1;
        B200LAP_PROF(S.sh->tr[TR_RECORDS] += total; S.sh->tr[TR_CYC_COLLECT_REPLAY] += sm_clock() - t0);
    }
}

template <typename CT>
__device__ __forceinline__ void replay_relax(SolverCtx<CT>& S, int hi_in, int wlo, int whi)
{
    const int lane = lane_id();
    const long long t0 = sm_clock();
    int hi = hi_in, fin = -1;
    for (int w0 = wlo; w0 <= whi; w0 += 32) {
        unsigned int bits = 0;
        if (w0 + lane <= whi) { bits = S.bitmap[w0 + lane]; S.bitmap[w0 + lane] = 0u; }
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
                    const int j = S.cols[k];
                    if (S.y[j] < 0) { fin = j; break; }
                    const int c2 = S.cols[hi];
                    S.cols[k] = c2; S.pos[c2] = k;
                    S.cols[hi] = j; S.pos[j] = hi;
                    ++hi;
                }
            }
        }
    }
    if (lane == 0) {
        S.sh->hi = hi; S.sh->final_j = fin;
        B200LAP_PROF(S.sh->tr[TR_RELAX_HITS] += hi - hi_in + (fin >= 0); S.sh->tr[TR_CYC_RELAX_REPLAY] += sm_clock() - t0);
    }
}

// ---- cluster mode: one CTA's slice of a relax step (same arithmetic and flags as the single-CTA body) ----------
template <typename CT>
__device__ __forceinline__ void relax_slice(SolverCtx<CT>& S, int i, int js, int hi, int sp)
{
    const int n = S.n, T = blockDim.x, tid = threadIdx.x;
    const int per = (n + S.nc - 1) / S.nc;
    const int j0 = S.rank * per, j1 = min(n, j0 + per);
    const CT* crow = S.C + (size_t)i * S.ld;
    // level and slack of the scanned column are fetched by every thread itself (one broadcast address each, in
    // flight together with the slice): the master posts the step without waiting for them
    const CT c_js = __ldg(crow + js);
    const double v_js = S.vg[js];
    const double level = S.d[js];        // js is a SCAN column: no relax step writes its distance
    const double slack = ((double)c_js - v_js) - level;
    int wmin_i = 0x7fffffff, wmax_i = -1;
    for (int base = j0 + tid; base < j1; base += 4 * T) {
        CT cr[4];
        int kq[4];
        double vq[4], dq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = base + q * T;
            cr[q] = j < j1 ? __ldg(crow + j) : (CT)0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = base + q * T;
            kq[q] = j < j1 ? S.pos[j] : -1;
            vq[q] = j < j1 ? S.vg[j] : 0.0;
            dq[q] = j < j1 ? S.d[j] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = base + q * T, k = kq[q];
            if (k >= hi) {
                const double cand = ((double)cr[q] - vq[q]) - slack;
                if (cand < dq[q]) {
                    S.d[j] = cand;
                    S.pred[j] = i;
                    if (cand == level) {
                        atomicOr(&S.bitmap[k >> 5], 1u << (k & 31));
                        wmin_i = min(wmin_i, k >> 5);
                        wmax_i = max(wmax_i, k >> 5);
                        if (dsm_add(S.msh + (unsigned)offsetof(SolverShared, nhit) + 4u * sp, 1) == 0) {
                            dsm_st(S.msh + (unsigned)offsetof(SolverShared, hit_k) + 4u * sp, k);
                            dsm_st(S.msh + (unsigned)offsetof(SolverShared, hit_j) + 4u * sp, j);
                        }
                    }
                }
            }
        }
    }
    if (wmax_i >= 0) {
        dsm_min(S.msh + (unsigned)offsetof(SolverShared, minw) + 4u * sp, wmin_i);
        dsm_max(S.msh + (unsigned)offsetof(SolverShared, maxw) + 4u * sp, wmax_i);
    }
}

// ---- cluster mode, batched: one CTA's slice of a relax step over up to kMaxScans SCAN columns at once.  The two cluster
// barriers of a step (and the L2 round trips of pos / v / d) are paid once per batch instead of once per scanned column.
// A thread relaxes its columns scan after scan in registers -- exactly the reference's sequential semantics per column
// (lapjv.cpp:178-213): the distance a later scan compares with is the one the earlier scans left, and a column hit by
// an earlier scan of the batch (cand == level: it moves to SCAN) is not touched by the later ones.  Hits are flagged per
// scan in a bitmap indexed by COLUMN (positions move while the earlier scans of the batch are replayed); the master
// turns every scan's column bits into position bits under the positions current at that moment and replays them in
// ascending position order with the single-scan machinery.  If a scan ends the path, the later scans of the batch
// never happened: their d / pred writes only touch TODO columns, which the next path re-initialises.
template <typename CT>
__device__ __forceinline__ void relax_slice_batch(SolverCtx<CT>& S, int K, const int* rows, const int* jss, int hi, int sp)
{
    const int n = S.n, T = blockDim.x, tid = threadIdx.x;
    const int per = (n + S.nc - 1) / S.nc;
    const int j0 = S.rank * per, j1 = min(n, j0 + per);
    const double level = S.d[jss[0]];          // every SCAN column sits at the current level
    const CT* crow[kMaxScans];
    double slack[kMaxScans];
#pragma unroll
    for (int sc = 0; sc < kMaxScans; ++sc) {
        const int r = sc < K ? rows[sc] : rows[0], js = sc < K ? jss[sc] : jss[0];
        crow[sc] = S.C + (size_t)r * S.ld;
        slack[sc] = ((double)__ldg(crow[sc] + js) - S.vg[js]) - level;
    }
    int wmin_i[kMaxScans], wmax_i[kMaxScans];
#pragma unroll
    for (int sc = 0; sc < kMaxScans; ++sc) { wmin_i[sc] = 0x7fffffff; wmax_i[sc] = -1; }
    for (int base = j0 + tid; base < j1; base += 2 * T) {
        CT cr[kMaxScans][2];
        int kq[2];
        double vq[2], dq[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = base + q * T;
#pragma unroll
            for (int sc = 0; sc < kMaxScans; ++sc) cr[sc][q] = (j < j1 && sc < K) ? __ldg(crow[sc] + j) : (CT)0;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = base + q * T;
            kq[q] = j < j1 ? S.pos[j] : -1;
            vq[q] = j < j1 ? S.vg[j] : 0.0;
            dq[q] = j < j1 ? S.d[j] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = base + q * T;
            if (kq[q] < hi) continue;
            double dj = dq[q];
            int pj = -1;
#pragma unroll
            for (int sc = 0; sc < kMaxScans; ++sc) {
                if (sc >= K) break;
                const double cand = ((double)cr[sc][q] - vq[q]) - slack[sc];
                if (cand < dj) {
                    dj = cand;
                    pj = rows[sc];
                    if (cand == level) {
                        atomicOr(&S.cbm[(size_t)sc * S.cbw + (j >> 5)], 1u << (j & 31));
                        wmin_i[sc] = min(wmin_i[sc], j >> 5);
                        wmax_i[sc] = max(wmax_i[sc], j >> 5);
                        if (dsm_add(S.msh + (unsigned)offsetof(SolverShared, bnhit) + 4u * (unsigned)(sp * kMaxScans + sc), 1) == 0)
                            dsm_st(S.msh + (unsigned)offsetof(SolverShared, bhit_j) + 4u * (unsigned)(sp * kMaxScans + sc), j);
                        break;                  // the column has left TODO for the later scans of the batch
                    }
                }
            }
            if (pj >= 0) { S.d[j] = dj; S.pred[j] = pj; }
        }
    }
#pragma unroll
    for (int sc = 0; sc < kMaxScans; ++sc) {
        if (wmax_i[sc] >= 0) {
            dsm_min(S.msh + (unsigned)offsetof(SolverShared, bminw) + 4u * (unsigned)(sp * kMaxScans + sc), wmin_i[sc]);
            dsm_max(S.msh + (unsigned)offsetof(SolverShared, bmaxw) + 4u * (unsigned)(sp * kMaxScans + sc), wmax_i[sc]);
        }
    }
}

template <typename CT>
__device__ void worker_loop(SolverCtx<CT>& S)
{
    for (;;) {
        cluster_sync_all();                                   // a command is posted
        const int op = dsm_ld(S.msh + (unsigned)offsetof(SolverShared, box_op));
        if (op == BOX_EXIT) { cluster_sync_all(); return; }   // the master's shared memory stays alive until everybody has read the command
        const int K = dsm_ld(S.msh + (unsigned)offsetof(SolverShared, box_k));
        const int hi = dsm_ld(S.msh + (unsigned)offsetof(SolverShared, box_hi)), sp = dsm_ld(S.msh + (unsigned)offsetof(SolverShared, box_sp));
        int rows[kMaxScans], jss[kMaxScans];
#pragma unroll
        for (int sc = 0; sc < kMaxScans; ++sc) {
            rows[sc] = sc < K ? dsm_ld(S.msh + (unsigned)offsetof(SolverShared, box_rows) + 4u * sc) : 0;
            jss[sc] = sc < K ? dsm_ld(S.msh + (unsigned)offsetof(SolverShared, box_jss) + 4u * sc) : 0;
        }
        relax_slice_batch(S, K, rows, jss, hi, sp);
        cluster_sync_all();                                   // slices done, flags and d/pred visible
    }
}

// cluster mode: the hits of one scan of a batch, flagged by COLUMN, become position bits under the CURRENT positions
// (warp 0; the words of the scan's range are cleared on the way).  Returns the position-word range through wlo / whi.
template <typename CT>
__device__ __forceinline__ void column_bits_to_positions(SolverCtx<CT>& S, int sc, int cw_lo, int cw_hi, int& wlo, int& whi)
{
    const int lane = lane_id();
    int mn = 0x7fffffff, mx = -1;
    unsigned int* cb = S.cbm + (size_t)sc * S.cbw;
    for (int w0 = cw_lo; w0 <= cw_hi; w0 += 32) {
        unsigned int bits = 0;
        if (w0 + lane <= cw_hi) { bits = cb[w0 + lane]; if (bits) cb[w0 + lane] = 0u; }
        while (bits) {
            const int bpos = __ffs((int)bits) - 1;
            bits &= bits - 1;
            const int k = S.pos[(w0 + lane) * 32 + bpos];
            atomicOr(&S.bitmap[k >> 5], 1u << (k & 31));
            mn = min(mn, k >> 5); mx = max(mx, k >> 5);
        }
    }
    __syncwarp();
    wlo = __reduce_min_sync(kFull, mn);
    whi = __reduce_max_sync(kFull, mx);
}

}  // namespace b200lap
#include "solver_path.cuh"
namespace b200lap {

// ---- one shortest augmenting path (find_path_dense) ---------------------------------------------
template <int MAXC, bool SMALLREG, typename CT>
__device__ int shortest_path(SolverCtx<CT>& S, int start_i)
{
    const int n = S.n, T = blockDim.x, tid = threadIdx.x;
    SolverShared* sh = S.sh;
    const CT* row0 = S.C + (size_t)start_i * S.ld;
    row_scan<MAXC>(row0, n, [&](int j, double c) {
        S.cols[j] = j;
        S.pos[j] = j;
        S.pred[j] = start_i;
        S.d[j] = c - S.v[j];
        if (S.vg != S.v) S.vg[j] = S.v[j];     // cluster mode: the workers' copy of v for this path
    });
    if (S.nc > 1)                              // cluster mode: scans discarded when the previous path ended may have left column bits
        for (int w = tid; w < kMaxScans * S.cbw; w += T) S.cbm[w] = 0u;
    __syncthreads();
    int lo = 0, hi = 0, n_ready = 0, final_j = -1;
    while (final_j < 0) {
        if (lo == hi) {
            // ---- level collect: positions [lo, n) in blocked ownership, prefix-min records flagged
            const long long tc0 = sm_clock();
            n_ready = lo;
            const int L = n - lo;
            const int chunk = (L + T - 1) / T;
            const int k0 = lo + tid * chunk;
            const int k1 = min(n, k0 + chunk);
            double lm = INFINITY;
            // the thread's chunk of distances is gathered ONCE, as independent loads, and kept in registers
            constexpr int kChunkRegs = MAXC > 0 ? MAXC : 1;
            double dch[kChunkRegs];
            if constexpr (MAXC > 0) {
                int cj[MAXC];
#pragma unroll
                for (int q = 0; q < MAXC; ++q) cj[q] = (q < chunk && k0 + q < k1) ? S.cols[k0 + q] : -1;
#pragma unroll
                for (int q = 0; q < MAXC; ++q) dch[q] = cj[q] >= 0 ? S.d[cj[q]] : INFINITY;
#pragma unroll
                for (int q = 0; q < MAXC; ++q) lm = dch[q] < lm ? dch[q] : lm;
            } else {
                for (int k = k0; k < k1; ++k) { const double t = S.d[S.cols[k]]; lm = t < lm ? t : lm; }
            }
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
            {
                const double t = lane_id() < warp_id() ? S.R.r->d[p][lane_id()] : INFINITY;
                const double wmin = warp_min_d(t);
                before = wmin < before ? wmin : before;
            }
            const int sp = S.step % 3;
            int wmin_i = 0x7fffffff, wmax_i = -1;
            double run = before;
            auto flag_record = [&](int k, double dj) {
                if (dj <= run) {
                    atomicOr(&S.bitmap[k >> 5], 1u << (k & 31));
                    wmin_i = min(wmin_i, k >> 5);
                    wmax_i = max(wmax_i, k >> 5);
                    run = dj;
                }
            };
            if constexpr (MAXC > 0) {
#pragma unroll
                for (int q = 0; q < MAXC; ++q)
                    if (q < chunk && k0 + q < k1) flag_record(k0 + q, dch[q]);
            } else {
                for (int k = k0; k < k1; ++k) flag_record(k, S.d[S.cols[k]]);
            }
            if (wmax_i >= 0) { atomicMin(&S.minw[sp], wmin_i); atomicMax(&S.maxw[sp], wmax_i); }
            __syncthreads();
            if (warp_id() == 0) {
                int wlo = S.minw[sp], whi = S.maxw[sp];
                if (lane_id() == 0) {
                    const int old_slot = (sp + 2) % 3;
                    S.minw[old_slot] = 0x7fffffff; S.maxw[old_slot] = -1; S.nhit[old_slot] = 0;
                    sh->tr[TR_COLLECT]++;
                }
                if (whi < 0) { wlo = lo >> 5; whi = wlo; }
                replay_collect(S, lo, wlo, whi);
            }
            __syncthreads();
            S.step++;
            hi = sh->hi;
            final_j = sh->final_j;
            B200LAP_PROF(if (tid == 0) sh->tr[TR_CYC_COLLECT] += sm_clock() - tc0);
        }
        // ---- relax from every SCAN column in turn (_scan_dense)
        const long long tr0 = sm_clock();
        while (final_j < 0 && lo != hi) {
            if (S.nc > 1) {
                // ---- cluster mode: up to kcap SCAN columns per step (relax_slice_batch), replayed scan by scan
                const int sp = S.step % 3;
                const int K = min(min(S.kcap, kMaxScans), hi - lo);
                if (tid < K) { const int js = S.cols[lo + tid]; sh->box_jss[tid] = js; sh->box_rows[tid] = S.y[js]; }
                if (tid == 0) { sh->box_op = BOX_RELAX; sh->box_k = K; sh->box_hi = hi; sh->box_sp = sp; }
                cluster_sync_all();
                relax_slice_batch(S, K, sh->box_rows, sh->box_jss, hi, sp);
                cluster_sync_all();
                S.step++;
                int consumed = 0;
                for (int sc = 0; sc < K && final_j < 0; ++sc) {
                    ++consumed;
                    const int nh = sh->bnhit[sp][sc];
                    if (nh == 0) continue;
                    if (nh == 1) {
                        if (tid == 0) {
                            const int j = sh->bhit_j[sp][sc];
                            S.cbm[(size_t)sc * S.cbw + (j >> 5)] = 0u;       // its only bit in that word range
                            const int k = S.pos[j];
                            int fin = -1, nhi = hi;
                            if (S.y[j] < 0) {
                                fin = j;
                            } else {
                                const int c2 = S.cols[hi];
                                S.cols[k] = c2; S.pos[c2] = k;
                                S.cols[hi] = j; S.pos[j] = hi;
                                nhi = hi + 1;
                            }
                            sh->hi = nhi; sh->final_j = fin;
                            B200LAP_PROF(sh->tr[TR_RELAX_HITS] += 1);
                        }
                    } else if (warp_id() == 0) {
                        int wlo, whi;
                        column_bits_to_positions(S, sc, sh->bminw[sp][sc], sh->bmaxw[sp][sc], wlo, whi);
                        replay_relax(S, hi, wlo, whi);
                    }
                    __syncthreads();
                    hi = sh->hi;
                    final_j = sh->final_j;
                }
                lo += consumed;
                // the slot is reset right after its use (collect steps advance the slot rotation too, but know nothing of the
                // per-scan arrays); the workers write it again only behind the next step's first cluster barrier
                __syncthreads();
                if (tid == 0) {
                    for (int sc = 0; sc < kMaxScans; ++sc) { sh->bnhit[sp][sc] = 0; sh->bminw[sp][sc] = 0x7fffffff; sh->bmaxw[sp][sc] = -1; }
                    sh->tr[TR_RELAX] += consumed;
                }
                continue;
            }
            const int js = S.cols[lo];
            const int i = S.y[js];
            const bool solo = S.nc == 1;             // cluster mode: every CTA fetches these three itself (relax_slice)
            const double level = solo ? S.d[js] : 0.0;
            ++lo;
            const CT* crow = S.C + (size_t)i * S.ld;
            const CT c_js = solo ? __ldg(crow + js) : (CT)0;        // issued together with the row loads below
            const double v_js = solo ? S.v[js] : 0.0;
            const int sp = S.step % 3;
            int wmin_i = 0x7fffffff, wmax_i = -1;
            auto relax_one = [&](int j, int k, double c, double vj, double dj, double slack) {
                if (k >= hi) {
                    const double cand = (c - vj) - slack;
                    if (cand < dj) {
                        S.d[j] = cand;
                        S.pred[j] = i;
                        if (cand == level) {
                            atomicOr(&S.bitmap[k >> 5], 1u << (k & 31));
                            wmin_i = min(wmin_i, k >> 5);
                            wmax_i = max(wmax_i, k >> 5);
                            if (atomicAdd(&S.nhit[sp], 1) == 0) { S.hit_k[sp] = k; S.hit_j[sp] = j; }
                        }
                    }
                }
            };
            if (S.nc > 1) {
                if (tid == 0) { sh->box_op = BOX_RELAX; sh->box_row = i; sh->box_js = js; sh->box_hi = hi; sh->box_sp = sp; }
                cluster_sync_all();
                relax_slice(S, i, js, hi, sp);
                cluster_sync_all();
            } else if constexpr (MAXC > 0) {
                // every load of the step (matrix row AND the thread's pos/v/d entries) is issued before the
                // first dependent instruction: the compiler cannot hoist shared loads over the stores of the
                // previous column by itself, and a column-at-a-time body costs ~700 cycles per column
                // The matrix-row loads (the long latency) are all in flight at once; the thread's pos/v/d entries
                // follow in groups of CH columns so that the 64-register variants (1024 threads) do not spill.
                constexpr int CH = SMALLREG ? (MAXC < 4 ? MAXC : 4) : (MAXC < 8 ? MAXC : 8);
                CT cr[MAXC];
#pragma unroll
                for (int q = 0; q < MAXC; ++q) {
                    const int j = tid + q * T;
                    cr[q] = j < n ? __ldg(crow + j) : (CT)0;
                }
                const double slack = ((double)c_js - v_js) - level;
#pragma unroll
                for (int q0 = 0; q0 < MAXC; q0 += CH) {
                    int kq[CH];
                    double vq[CH], dq[CH];
#pragma unroll
                    for (int q = 0; q < CH; ++q) {
                        const int j = tid + (q0 + q) * T;
                        kq[q] = j < n ? S.pos[j] : -1;
                        vq[q] = j < n ? S.v[j] : 0.0;
                        dq[q] = j < n ? S.d[j] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < CH; ++q) relax_one(tid + (q0 + q) * T, kq[q], (double)cr[q0 + q], vq[q], dq[q], slack);
                }
            } else {
                const double slack = ((double)c_js - v_js) - level;
                for (int j = tid; j < n; j += T) relax_one(j, S.pos[j], (double)crow[j], S.v[j], S.d[j], slack);
            }
            if (wmax_i >= 0) { atomicMin(&S.minw[sp], wmin_i); atomicMax(&S.maxw[sp], wmax_i); }
            if (S.nc == 1) __syncthreads();       // (cluster mode: the closing cluster barrier already did this)
            S.step++;
            // one round trip for the whole slot (in cluster mode these live in the global mailbox)
            const int whi = S.maxw[sp], nh = S.nhit[sp], hk = S.hit_k[sp], hj = S.hit_j[sp];
            if (tid == 0) {
                // the slot used one step ago has been read by everyone (they all passed this barrier);
                // it is next written two steps from now, after another barrier
                const int old_slot = (sp + 2) % 3;
                S.minw[old_slot] = 0x7fffffff; S.maxw[old_slot] = -1; S.nhit[old_slot] = 0;
                sh->tr[TR_RELAX]++;
            }
            if (whi >= 0) {
                if (nh == 1) {
                    // the common case, one hit: its (position, column) was published by the thread that found it, so the
                    // swap needs no bitmap walk (three dependent shared loads instead of about twelve)
                    if (tid == 0) {
                        const int k = hk, j = hj;
                        S.bitmap[k >> 5] = 0u;
                        int fin = -1, nhi = hi;
                        if (S.y[j] < 0) {
                            fin = j;
                        } else {
                            const int c2 = S.cols[hi];
                            S.cols[k] = c2; S.pos[c2] = k;
                            S.cols[hi] = j; S.pos[j] = hi;
                            nhi = hi + 1;
                        }
                        sh->hi = nhi; sh->final_j = fin;
                        B200LAP_PROF(sh->tr[TR_RELAX_HITS] += 1);
                    }
                } else if (warp_id() == 0) {
                    const int wlo = S.minw[sp];
                    replay_relax(S, hi, wlo, whi);
                }
                __syncthreads();
                hi = sh->hi;
                final_j = sh->final_j;
            }
        }
        B200LAP_PROF(if (tid == 0) sh->tr[TR_CYC_RELAX] += sm_clock() - tr0);
    }
    // ---- dual update of the READY columns (lapjv.cpp:270-276); lo of the caller == n_ready
    const double level = S.d[S.cols[n_ready]];
    for (int j = tid; j < n; j += T)
        if (S.pos[j] < n_ready) S.v[j] += S.d[j] - level;
    return final_j;
}

template <int MAXC, bool SMALLREG, typename CT>
__device__ void augment_all(SolverCtx<CT>& S, int n_free)
{
    for (int f = 0; f < n_free; ++f) {
        const int root = S.free_rows[f];
        int col;
        if constexpr (MAXC > 0 && MAXC % VecOf<CT>::V == 0) {
            col = S.regpath ? shortest_path_reg<MAXC>(S, root) : shortest_path<MAXC, SMALLREG>(S, root);
        } else {
            col = shortest_path<MAXC, SMALLREG>(S, root);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            S.sh->tr[TR_PATHS]++;
            int r;
            do {
                r = S.pred[col];
                S.y[col] = r;
                const int prev = S.x[r];
                S.x[r] = col;
                col = prev;
            } while (r != root);
        }
        __syncthreads();
    }
}

// ---- ascending list of rows with x[i] < 0 (warp 0, ballot compaction) -> count -------------------
template <typename CT>
__device__ int collect_free_rows(SolverCtx<CT>& S)
{
    __syncthreads();
    if (warp_id() == 0) {
        int cnt = 0;
        for (int i0 = 0; i0 < S.n; i0 += 32) {
            const int i = i0 + lane_id();
            const bool is_free = i < S.n && S.x[i] < 0;
            const unsigned int m = __ballot_sync(kFull, is_free);
            if (is_free) S.free_rows[cnt + __popc(m & ((1u << lane_id()) - 1u))] = i;
            cnt += __popc(m);
        }
        if (lane_id() == 0) S.sh->aux = cnt;
    }
    __syncthreads();
    return S.sh->aux;
}

// ---- cold solve: column reduction + reduction transfer (_ccrrt_dense) ----------------------------
template <int MAXC, typename CT>
__device__ int col_reduce(SolverCtx<CT>& S, const CT* colmin, const int* colarg)
{
    const int n = S.n, T = blockDim.x, tid = threadIdx.x;
    int* cnt = S.pred;   // scratch: how many columns chose each row
    for (int j = tid; j < n; j += T) {
        const double cm = (double)colmin[j];
        const bool hit = cm < B200LAP_LARGE;
        S.v[j] = hit ? cm : B200LAP_LARGE;
        S.y[j] = hit ? colarg[j] : 0;
        S.x[j] = -1;
        cnt[j] = 0;
    }
    __syncthreads();
    // reverse column sweep == every row keeps its HIGHEST column; rows chosen twice are not unique
    for (int j = tid; j < n; j += T) {
        const int i = S.y[j];
        atomicMax(&S.x[i], j);
        atomicAdd(&cnt[i], 1);
    }
    __syncthreads();
    for (int j = tid; j < n; j += T)
        if (S.x[S.y[j]] != j) S.y[j] = -1;
    const int n_free = collect_free_rows(S);
    // reduction transfer, rows ascending, sequential through v (lapjv.cpp:52-68)
    for (int i = 0; i < n; ++i) {
        const int own = S.x[i];
        if (own < 0 || cnt[i] != 1) continue;     // uniform: shared state read after a barrier
        const CT* crow = S.C + (size_t)i * S.ld;
        double m = B200LAP_LARGE;
        row_scan<MAXC>(crow, n, [&](int j, double c) {
            if (j != own) {
                const double red = c - S.v[j];
                m = red < m ? red : m;
            }
        });
        m = red_min_d(S.R, m);
        if (tid == 0) S.v[own] -= m;
        __syncthreads();
    }
    return n_free;
}

// ---- cold solve: one augmenting-row-reduction pass (_carr_dense) -----------------------------------
template <int MAXC, typename CT>
__device__ int arr_pass(SolverCtx<CT>& S, int n_free)
{
    const int n = S.n, T = blockDim.x, tid = threadIdx.x;
    SolverShared* sh = S.sh;
    unsigned int cursor = 0, steps = 0;
    int deferred = 0;
    __syncthreads();
    while (cursor < (unsigned int)n_free) {
        ++steps;
        const long long ta0 = sm_clock();
        const int r = S.free_rows[cursor++];
        const CT* crow = S.C + (size_t)r * S.ld;
        const double c0 = (double)__ldg(crow) - S.v[0];
        Top2 t;
        top2_init(t);
        int k1, k2;
        double b1, b2;
        // One pass gathers both candidate sets; which one applies depends on c_0 (uniform).
        //   c_0 <  LARGE: lexicographic top-2 of {(c_0,0)} U {(c_j,j): j>=1, c_j < LARGE}   
        //   c_0 >= LARGE: entries before the first j >= 1 below LARGE are skipped, then no filter at all
        Top2 tf;            // filtered set
        top2_init(tf);
        int f = 0x7fffffff;
        if constexpr (MAXC >= 16) {
            // (measured: 8 x 8192 cold fallback 5464 -> 5042 ms with 16 entries per thread; with 8 entries per thread the
            // running form is 6 % faster, so shorter rows keep it)
            // The lexicographic (value, index) top-2 of the thread's entries by a TOURNAMENT instead of a running update:
            // the running form (generic branch below) carries a1 / a2 from entry to entry -- ncu of the cold n = 4096 solve
            // (profiles/r02_ncu_cold_solve_4096.txt) shows 33 instructions and ~100 cycles per entry with the v[j] load and
            // its dependent DADD inside each entry's branch.  Here every load and every subtraction is independent, a
            // filtered entry becomes the empty slot (+inf, INT_MAX), and log2(MAXC) levels of merges follow.  Entries are in
            // increasing column order and a merge prefers its left operand on ties, so the result is the same pair.
            double ra[MAXC];
            int ri[MAXC];
#pragma unroll
            for (int q = 0; q < MAXC; ++q) {
                const int j = tid + q * T;
                ra[q] = j < n ? (double)__ldg(crow + j) : INFINITY;
            }
#pragma unroll
            for (int q = 0; q < MAXC; ++q) {
                const int j = tid + q * T;
                const double red = ra[q] - (j < n ? S.v[j] : 0.0);
                const bool ok = j < n && j != 0 && red < B200LAP_LARGE;
                ra[q] = ok ? red : INFINITY;
                ri[q] = ok ? j : 0x7fffffff;
                f = ok ? min(f, j) : f;
            }
            Top2 tt[(MAXC + 1) / 2];
#pragma unroll
            for (int q = 0; q < MAXC; q += 2) {
                Top2& o = tt[q / 2];
                if (q + 1 < MAXC) {
                    const bool p = ra[q + 1] < ra[q];
                    o.a1 = p ? ra[q + 1] : ra[q]; o.i1 = p ? ri[q + 1] : ri[q];
                    o.a2 = p ? ra[q] : ra[q + 1]; o.i2 = p ? ri[q] : ri[q + 1];
                } else {
                    o.a1 = ra[q]; o.i1 = ri[q]; o.a2 = INFINITY; o.i2 = 0x7fffffff;
                }
            }
#pragma unroll
            for (int w = 1; w < (MAXC + 1) / 2; w *= 2) {
#pragma unroll
                for (int q = 0; q + w < (MAXC + 1) / 2; q += 2 * w) {
                    Top2& A = tt[q];
                    const Top2 B = tt[q + w];
                    const bool p = B.a1 < A.a1;                      // the right operand wins only when strictly smaller
                    const double xa = p ? A.a1 : A.a2; const int xi = p ? A.i1 : A.i2;
                    const double ya = p ? B.a2 : B.a1; const int yi = p ? B.i2 : B.i1;
                    const bool s2 = ya < xa;
                    A.a1 = p ? B.a1 : A.a1; A.i1 = p ? B.i1 : A.i1;
                    A.a2 = s2 ? ya : xa;    A.i2 = s2 ? yi : xi;
                }
            }
            tf = tt[0];
        } else {
            row_scan<MAXC>(crow, n, [&](int j, double c) {
                if (j != 0) {
                    const double red = c - S.v[j];
                    if (red < B200LAP_LARGE) { top2_push_inc(tf, red, j); f = min(f, j); }
                }
            });
        }
        if (c0 < B200LAP_LARGE) {
            if (tid == 0) top2_push(tf, c0, 0);
            t = block_top2(S.R, tf);
        } else {
            f = red_min_i(S.R, f);
            if (f != 0x7fffffff) {
                if (tid == 0) top2_push(t, c0, 0);
                row_scan<MAXC>(crow, n, [&](int j, double c) {
                    if (j >= f) top2_push_inc(t, c - S.v[j], j);
                });
            } else if (tid == 0) {
                top2_push(t, c0, 0);
            }
            t = block_top2(S.R, t);
        }
        k1 = t.i1; b1 = t.a1;
        if (t.i2 == 0x7fffffff) { k2 = -1; b2 = B200LAP_LARGE; } else { k2 = t.i2; b2 = t.a2; }
        const long long ta1 = sm_clock();
        if (tid == 0) {
            sh->tr[TR_ARR]++;
            B200LAP_PROF(sh->tr[TR_CYC_ARR_SCAN] += ta1 - ta0);
            int owner = S.y[k1];
            const double lowered = S.v[k1] - (b2 - b1);
            const bool does_lower = lowered < S.v[k1];
            if (steps < cursor * (unsigned int)n) {
                if (does_lower) {
                    S.v[k1] = lowered;
                } else if (owner >= 0 && k2 >= 0) {
                    k1 = k2;
                    owner = S.y[k2];
                }
                if (owner >= 0) {
                    if (does_lower) S.free_rows[--cursor] = owner;
                    else S.free_rows[deferred++] = owner;
                }
            } else if (owner >= 0) {
                S.free_rows[deferred++] = owner;
            }
            S.x[r] = k1;
            S.y[k1] = r;
            sh->cursor = cursor;
            sh->deferred = (unsigned int)deferred;
        }
        __syncthreads();
        B200LAP_PROF(if (tid == 0) sh->tr[TR_CYC_ARR_SERIAL] += sm_clock() - ta1);
        cursor = sh->cursor;
        deferred = (int)sh->deferred;
    }
    return deferred;
}

// returns the number of rows still free (S.free_rows[0..left)): the caller augments them
template <int MAXC, typename CT>
__device__ int cold_solve(SolverCtx<CT>& S, const CT* colmin, const int* colarg)
{
    int left = col_reduce<MAXC>(S, colmin, colarg);
    if (threadIdx.x == 0) S.sh->tr[TR_FREE_CR] = left;
    for (int pass = 0; left > 0 && pass < 2; ++pass) left = arr_pass<MAXC>(S, left);
    __syncthreads();
    return left;
}

// ---- the persistent per-instance kernel -----------------------------------------------------------
template <typename CT, int MAXC, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_solve(SolveArgs<CT> a)
{
    B200LAP_DYN_SMEM(dyn);
    __shared__ SolverShared sh;
    const int b = a.cluster > 1 ? (int)(blockIdx.x / (unsigned)a.cluster) : (int)blockIdx.x;
    const int n = a.n, T = blockDim.x, tid = threadIdx.x;
    SolverCtx<CT> S;
    S.C = a.C + (size_t)b * a.inst_stride;
    S.ld = a.ld;
    S.n = n;
    {
        unsigned char* sbase = dyn;
        unsigned char* gbase = a.gws ? a.gws + (size_t)b * a.gws_stride : nullptr;
        auto place = [&](int arr) -> void* {
            const size_t bytes = state_array_bytes(arr, n);
            unsigned char* p;
            if (a.smem_mask & (1 << arr)) { p = sbase; sbase += bytes; }
            else { p = gbase; gbase += bytes; }
            return p;
        };
        // same order as solver_place_state so shared offsets match the host's byte count
        S.bitmap = (unsigned int*)place(ST_BITMAP);
        S.d = (double*)place(ST_D);
        S.pos = (int*)place(ST_POS);
        S.v = (double*)place(ST_V);
        S.y = (int*)place(ST_Y);
        S.cols = (int*)place(ST_COLS);
        S.pred = (int*)place(ST_PRED);
        S.x = (int*)place(ST_X);
        S.free_rows = (int*)place(ST_FREE);
        S.srow = (int*)place(ST_SROW);
        // the workspace stride covers every array, so whatever was placed in shared memory leaves room for the mirror
        S.vg = (a.cluster > 1 && (a.smem_mask & (1 << ST_V))) ? (double*)gbase : S.v;
    }
    S.sh = &sh;
    S.R.r = &sh.red;
    S.R.r2 = &sh.red2;
    S.R.par = 0;
    S.step = 0;
    S.nc = a.cluster > 1 ? a.cluster : 1;
    S.rank = a.cluster > 1 ? (int)(blockIdx.x % (unsigned)a.cluster) : 0;
    S.msh = S.nc > 1 ? dsm_map(&sh, 0u) : 0u;
    // cluster mode: the per-scan column bitmaps of the batched relax step sit at the tail of the instance's workspace stride
    S.cbw = ((n + 31) / 32 + 4 + 3) & ~3;
    S.cbm = (S.nc > 1 && a.gws) ? (unsigned int*)(a.gws + (size_t)(b + 1) * a.gws_stride) - (size_t)kMaxScans * S.cbw : nullptr;
    S.regpath = (a.regpath && S.nc == 1 && a.smem_mask == (1 << ST_COUNT) - 1) ? 1 : 0;
    S.kcap = a.kcap > 0 ? a.kcap : kMaxScans;
    S.pipe = a.pipe;
    S.minw = sh.minw; S.maxw = sh.maxw; S.nhit = sh.nhit; S.hit_k = sh.hit_k; S.hit_j = sh.hit_j;
    if (S.rank != 0) { worker_loop(S); return; }
    if (tid == 0) {
        sh.s_cnt = 0;
        for (int q = 0; q < 3; ++q) {
            S.minw[q] = 0x7fffffff; S.nhit[q] = 0; S.maxw[q] = -1;
            for (int s2 = 0; s2 < kMaxScans; ++s2) { sh.nh[q][s2] = 0; sh.bnhit[q][s2] = 0; sh.bminw[q][s2] = 0x7fffffff; sh.bmaxw[q][s2] = -1; }
        }
        for (int q = 0; q < kTraceWords; ++q) sh.tr[q] = 0;
    }
    for (int w = tid; w < (n + 31) / 32 + 4; w += T) S.bitmap[w] = 0u;
    const CT* colmin = a.colmin + (size_t)b * n;
    const int* colarg = a.colarg + (size_t)b * n;
    int rc = 0;
    const long long t_start = sm_clock();
    __syncthreads();

    int n_aug = 0;          // rows left for the augmentation phase (one call site for all three ways to get here)
    if (a.mode == 1) {
        n_aug = cold_solve<MAXC>(S, colmin, colarg);
    } else {
        const double eps = a.eps;
        const double tol = eps > 1e-9 ? eps : 1e-9;
        const double* us = a.u_seed + (size_t)b * n;
        const double* vs = a.v_seed + (size_t)b * n;
        int* tl = a.tight_cols + (size_t)b * n * kTightCap;
        int* tc = a.tight_cnt + (size_t)b * n;
        double* u = S.d;   // row potentials live in d[] until the first path search
        const FrontFlags fl = a.flags[b];
        unsigned long long tight = fl.total_tight;
        int infeasible = fl.infeasible;
        for (int j = tid; j < n; j += T) { S.v[j] = vs[j]; S.x[j] = -1; S.y[j] = -1; }
        if (!fl.any_viol) {
            const double* ut = a.u_tight + (size_t)b * n;
            for (int j = tid; j < n; j += T) u[j] = ut[j];
            __syncthreads();
        } else {
            // projection fired somewhere: redo the whole front end in reference order
            __syncthreads();
            long long fired = 0;
            for (int i = 0; i < n; ++i) {
                double ui = us[i];
                fired += project_row(S.C + (size_t)i * S.ld, n, S.v, ui, eps, S.R);
                if (tid == 0) u[i] = ui;
            }
            __syncthreads();
            tight = 0;
            infeasible = 0;
            for (int i = 0; i < n; ++i) {
                double ut;
                int cnt;
                front_row_generic(S.C + (size_t)i * S.ld, n, S.v, u[i], eps, tol, S.R, &sh.s_cnt, sh.s_list,
                                  tl + (size_t)i * kTightCap, &ut, &cnt, &infeasible);
                if (tid == 0) { u[i] = ut; tc[i] = cnt; }
                tight += (unsigned long long)cnt;
            }
            if (tid == 0) sh.tr[TR_PROJ] = fired;
            __threadfence_block();
            __syncthreads();
        }
        if (infeasible) {
            rc = -3;
        } else {
            // ---- greedy first-fit over the tight lists (lapjv_seeded.cpp:76-93)
            int i = 0;
            while (i < n) {
                if (tid == 0) {
                    int matched = 0;
                    for (; i < n; ++i) {
                        const int c = tc[i];
                        if (c > kTightCap) break;          // list truncated: cooperative rescan below
                        const int* li = tl + (size_t)i * kTightCap;
                        for (int q = 0; q < c; ++q) {
                            const int j = li[q];
                            if (S.y[j] < 0) { S.x[i] = j; S.y[j] = i; ++matched; break; }
                        }
                    }
                    sh.next_row = i;
                    sh.tr[TR_GREEDY] += matched;
                }
                __syncthreads();
                i = sh.next_row;
                if (i < n) {
                    const CT* crow = S.C + (size_t)i * S.ld;
                    const double ui = u[i];
                    int first = 0x7fffffff;
                    row_scan<MAXC>(crow, n, [&](int j, double c) {
                        if (S.y[j] < 0 && fabs((c - ui) - S.v[j]) <= tol) first = min(first, j);
                    });
                    first = red_min_i(S.R, first);
                    if (tid == 0 && first != 0x7fffffff) { S.x[i] = first; S.y[first] = i; sh.tr[TR_GREEDY]++; }
                    ++i;
                    __syncthreads();
                }
            }
            const int n_free = collect_free_rows(S);
            if (tid == 0) sh.tr[TR_TIGHT] = (long long)tight;
            if ((double)(long long)tight < 1.2 * n) {
                if (tid == 0) sh.tr[TR_FALLBACK] = 1;
                __syncthreads();
                n_aug = cold_solve<MAXC>(S, colmin, colarg);
            } else if (n_free > 0) {
                // ---- micro-ARR (lapjv_seeded.cpp:136-159); "j1 in free_cols" == y[j1] < 0 after greedy
                for (int f = 0; f < n_free; ++f) {
                    const int r = S.free_rows[f];
                    const CT* crow = S.C + (size_t)r * S.ld;
                    const double ur = u[r];
                    Top2 t;
                    top2_init(t);
                    row_scan<MAXC>(crow, n, [&](int j, double c) { top2_push_inc(t, (c - ur) - S.v[j], j); });
                    t = block_top2(S.R, t);
                    if (tid == 0 && t.i1 != 0x7fffffff && t.a2 - t.a1 > tol && S.y[t.i1] < 0) {
                        S.v[t.i1] += t.a2 - t.a1;
                        sh.tr[TR_MICRO]++;
                    }
                    __syncthreads();
                }
                n_aug = n_free;
            }
        }
    }
    __syncthreads();
    if (n_aug > 0) augment_all<MAXC, (MAXT > 512)>(S, n_aug);
    __syncthreads();
    if (rc == 0) {
        for (int j = tid; j < n; j += T) {
            a.x[(size_t)b * n + j] = S.x[j];
            a.y[(size_t)b * n + j] = S.y[j];
            if (a.v_out) a.v_out[(size_t)b * n + j] = S.v[j];
        }
    }
    if (S.nc > 1) {
        if (tid == 0) sh.box_op = BOX_EXIT;
        cluster_sync_all();                 // the workers read the command ...
        cluster_sync_all();                 // ... and only then may the master's shared memory go away
    }
    if (tid == 0) {
        a.rc[b] = rc;
        sh.tr[TR_RC] = rc;
        B200LAP_PROF(sh.tr[TR_CYC_TOTAL] = sm_clock() - t_start);
        if (a.trace)
            for (int q = 0; q < kTraceWords; ++q) a.trace[(size_t)b * kTraceWords + q] = sh.tr[q];
    }
}

}  // namespace b200lap
