// features_smem.cuh -- the 21-D row features + per-row top-k for binary32 storage, built to run at the
// HBM roofline: persistent CTAs, rows staged in shared memory by the bulk async-copy engine, two fused
// passes over the on-chip row.
//
// Reference: gnn/features.py:161-243 (compute_row_features), :21-31 (_positional_encodings);
// gnn/one_gnn.py:143-147 (the top-k VALUES of cost - u_pre; a per-row constant shift is monotone, so
// the k smallest raw entries are selected here).  Definitions: SURVEY.md App. B.
//
// Why this shape.  One read of C is 4 bytes per entry; at 6.5 TB/s over 148 SMs that leaves ~22 issue
// slots per 32 entries and per SM sub-partition, so the kernel is INSTRUCTION bound and the design
// counts instructions per entry:
//   * a CTA (64..512 threads) owns one row at a time and loops over rows (persistent grid);  the row
//     arrives in shared memory by ONE cp.async.bulk (SASS UBLKCP) tracked by an mbarrier, issued for
//     the next row while the current one is processed (two buffers) or overlapped with the other
//     resident CTAs of the SM (one buffer);  no per-thread global loads of C at all.
//   * exact order statistics (median, MAD) cost ONE compare-and-compact per entry:  a strided sample
//     of the row (512..2048 keys) is histogrammed (256 value-linear bins, shared atomics -- cheap, the
//     sample is small) to bracket the wanted ranks by two sample order statistics [L, H];  the fused
//     pass then counts the entries below L (sign bit of c - L, one LEA.HI) and appends the ~10% of
//     entries inside the bracket to a THREAD-PRIVATE list (predicated STS + IADD, no atomics, no
//     divergence);  the exact rank is then resolved on the list alone (histogram levels, then a
//     <= 32-key warp ranking).  A bracket that misses (probability ~1e-3 per row), a list that
//     overflows or a tie-heavy row falls back to exact histogram levels over the whole on-chip row
//     -- slower, never wrong.  Rows whose bracket collapses to one value (sparse family: 70% of a row
//     is the 1e6 fill) run a count-only "tie" variant of the pass.
//   * pass 1 (min, max, sum, is_col_best, median bracket) and pass 2 (variance, exp sums, near-best
//     count, top-k candidates below the 16th smallest per-thread minimum, MAD bracket) are the only
//     sweeps over the row;  everything else touches O(sample) or O(list) keys.
//   * per-thread partials are binary32 (<= 128 entries each), cross-thread combination uses warp
//     redux on order-preserving integer images for min/max/counts and shuffles for sums;  scalar
//     finishing math is done once per row by one lane.
// Tolerance (tests): |delta| <= 1e-4 |ref| + 1e-7 per feature; the selected order statistics are exact.
#pragma once
#include "features.cuh"

namespace b200lap {

constexpr int kSelBins = 256;
constexpr int kCandCap = 256;
constexpr int kTinyCap = 32;
constexpr int kRedSlots = 8;

constexpr int kTinyCap2 = 64;

struct __align__(16) FeatSmemFixed {
    uint64_t mbar[2];
    int hist[2][kSelBins];
    float cand[kCandCap];
    float tiny[kTinyCap2];
    float sorted[kTopKMax];
    unsigned int red[2][kRedSlots][32];   // block_reduce scratch (fall-back paths)
    // per-row cells combined with shared atomics by lane 0 of every warp; one set per row parity
    unsigned int cell_u[2][8];            // order-preserving images: 0 sample min, 1 sample max, 2 row min, 3 row max, 4 top-k bound
    int cell_i[2][12];                    // 0 notbest 1 below 2 inside 3 overflow | 4 below 5 inside 6 overflow (MAD) | 7 near 8 ones
    float psum[2][6][32];                 // per-warp partial sums: 0 sum 1 sum z^2 2 sum exp 3 sum exp*z 4 exact ssq 5 small exp
    unsigned int cell_h[2];               // min / max of a heavy bracket (reset by the leader)
    float bc_f[14];                       // leader -> CTA broadcasts
    int bc_i[16];
    int ncand, ntiny;
};
static_assert(sizeof(FeatSmemFixed) % 16 == 0, "row buffers follow the fixed block and need 16-byte alignment");

struct FeatSmemArgs {
    const float* C;
    long long inst_stride;
    int ld, n, batch, topk;
    const float* colmin;   // [batch][n]
    float* feat;           // [batch][n][21]
    float* topv;           // [batch][n][topk] or null
    const float* posenc;   // [n][8]  (sin, cos) x (1, 2, 4, 8), gnn/features.py:21-31
    int nbuf;              // row buffers per CTA (1 or 2)
    int row_floats;        // floats per row buffer (multiple of 4)
    int nsamp;             // sample size (<= n)
    int delta;             // half width of the sample-rank bracket
    int kcap;              // thread-private list capacity (multiple of 4, or >= entries per thread)
    int use_bulk;          // rows arrive by cp.async.bulk (needs 16-byte aligned rows)
    const int* rows_in;        // optional indirection: process rows rows_in[0 .. *rows_in_count) (global row ids b*n + row)
    const int* rows_in_count;
    int* redo_list;            // warp kernel: rows handed to the CTA kernel's exact fall-backs
    int* redo_count;
    int torch_mode = 0;        // near-best threshold as compute_row_features_torch evaluates it: fl32(row_min * 1.1f) (gnn/features.py:316)
};

// c <= mn * 1.1 in binary64 (NumPy definition)  <=>  c <= round_down(mn * 1.1); the torch definition multiplies in binary32
__device__ __forceinline__ float near_threshold(float mn, int torch_mode) {
    return torch_mode ? mn * 1.1f : round_down_to<float>((double)mn * 1.1);
}

// ---- order-preserving integer image of a binary32 value (for warp redux min / max) -----------------
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return u ^ ((unsigned)((int)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) { return __uint_as_float(o ^ (((o >> 31) - 1u) | 0x80000000u)); }

enum { RK_FMIN = 0, RK_FMAX = 1, RK_FSUM = 2, RK_IADD = 3, RK_IMAX = 4 };

__device__ __forceinline__ unsigned red_warp(int kind, unsigned v) {
    switch (kind) {
        case RK_FMIN: return __float_as_uint(ord2f(__reduce_min_sync(kFull, f2ord(__uint_as_float(v)))));
        case RK_FMAX: return __float_as_uint(ord2f(__reduce_max_sync(kFull, f2ord(__uint_as_float(v)))));
        case RK_FSUM: return __float_as_uint(warp_sum_f(__uint_as_float(v)));
        case RK_IADD: return __reduce_add_sync(kFull, v);
        default: return __reduce_max_sync(kFull, v);
    }
}
__device__ __forceinline__ unsigned red_identity_u(int kind) {
    switch (kind) {
        case RK_FMIN: return 0x7f800000u;
        case RK_FMAX: return 0xff800000u;
        default: return 0u;
    }
}

// Up to kRedSlots independent block reductions behind ONE barrier; every warp finishes redundantly,
// so the results are uniform without a second barrier.  Slot arrays alternate with `par`.
template <int... KS>
__device__ __forceinline__ void block_reduce(FeatSmemFixed& F, int& par, unsigned (&v)[sizeof...(KS)]) {
    constexpr int N = sizeof...(KS);
    constexpr int kinds[N] = {KS...};
    static_assert(N <= kRedSlots, "too many reduction slots");
    par ^= 1;
    const int l = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
#pragma unroll
    for (int q = 0; q < N; ++q) v[q] = red_warp(kinds[q], v[q]);
    if (l == 0) {
#pragma unroll
        for (int q = 0; q < N; ++q) F.red[par][q][w] = v[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < N; ++q) v[q] = red_warp(kinds[q], l < nw ? F.red[par][q][l] : red_identity_u(kinds[q]));
}

__device__ __forceinline__ int sel_bin256(float x, float lo, float scale) {
    const float t = (x - lo) * scale;
    return t >= (float)(kSelBins - 1) ? kSelBins - 1 : (t > 0.0f ? (int)t : 0);
}

// same for a key known to lie in [lo, hi] (no clamping from below, no NaN care)
__device__ __forceinline__ int sel_bin256_in(float x, float lo, float scale) {
    const int b = (int)((x - lo) * scale);
    return b > kSelBins - 1 ? kSelBins - 1 : b;
}

// which of the 256 bins holds the member of (0-based) rank `rel`; every warp computes it redundantly
__device__ __forceinline__ void find_bin(const int* h, int rel, int& b, int& bbefore, int& bcnt) {
    const int l = lane_id();
    const int4 a = reinterpret_cast<const int4*>(h)[2 * l];
    const int4 d = reinterpret_cast<const int4*>(h)[2 * l + 1];
    const int c[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
    int tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) tot += c[q];
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (l >= o) incl += t;
    }
    int run = incl - tot;
    const unsigned m = __ballot_sync(kFull, rel >= run && rel < incl);
    const int src = m ? __ffs((int)m) - 1 : 31;
    int bi = 8 * l + 7, bb = run, bc = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (rel >= run && rel < run + c[q]) { bi = 8 * l + q; bb = run; bc = c[q]; }
        run += c[q];
    }
    b = __shfl_sync(kFull, bi, src);
    bbefore = __shfl_sync(kFull, bb, src);
    bcnt = __shfl_sync(kFull, bc, src);
}

// members = keys in [lo, hi];  before = keys below lo;  inside = members
struct SelState { float lo, hi; int before, inside; };

// One exact histogram level: narrows the state to the bin holding absolute rank t.  Needs lo < hi,
// F.hist[hpar] zero on entry; leaves F.hist[hpar ^ 1] zero and flips hpar.  Two barriers.
template <typename Each>
__device__ __forceinline__ void narrow_level(FeatSmemFixed& F, int& par, int& hpar, Each each, SelState& s, int t) {
    const int T = blockDim.x, tid = threadIdx.x;
    const float lo = s.lo, hi = s.hi, scale = (float)kSelBins / (hi - lo);
    int* h = F.hist[hpar];
    each([&](float x) { if (x >= lo && x <= hi) atomicAdd(&h[sel_bin256(x, lo, scale)], 1); });
    for (int i = tid; i < kSelBins; i += T) F.hist[hpar ^ 1][i] = 0;
    __syncthreads();
    int b, bb, bc;
    find_bin(h, t - s.before, b, bb, bc);
    float mn = INFINITY, mx = -INFINITY;
    each([&](float x) {
        if (x >= lo && x <= hi && sel_bin256(x, lo, scale) == b) { mn = fminf(mn, x); mx = fmaxf(mx, x); }
    });
    unsigned r[2] = {__float_as_uint(mn), __float_as_uint(mx)};
    block_reduce<RK_FMIN, RK_FMAX>(F, par, r);
    s.lo = __uint_as_float(r[0]); s.hi = __uint_as_float(r[1]);
    s.before += bb; s.inside = bc;
    hpar ^= 1;
}

// exact key of absolute rank t.  `exact`: lo / hi are attained by members (true after any level)
template <typename Each>
__device__ __forceinline__ float select_rank(FeatSmemFixed& F, int& par, int& hpar, Each each, SelState s, int t, bool exact) {
    while (true) {
        if (exact) {
            if (!(s.lo < s.hi) || s.inside <= 1) return s.lo;
            if (s.inside == 2) return t == s.before ? s.lo : s.hi;
        }
        if (s.inside <= kTinyCap) {
            // gather the members and rank them inside every warp
            const float lo = s.lo, hi = s.hi;
            each([&](float x) {
                if (x >= lo && x <= hi) { const int p = atomicAdd(&F.ntiny, 1); if (p < kTinyCap) F.tiny[p] = x; }
            });
            __syncthreads();
            const int l = lane_id(), cnt = s.inside, want = t - s.before;
            const float x = l < cnt ? F.tiny[l] : INFINITY;
            int rank = 0;
            for (int q = 0; q < cnt; ++q) {
                const float o = __shfl_sync(kFull, x, q);
                rank += (o < x) || (o == x && q < l);
            }
            const unsigned hit = __ballot_sync(kFull, l < cnt && rank == want);
            const float res = __shfl_sync(kFull, x, hit ? __ffs((int)hit) - 1 : 0);
            __syncthreads();                 // every warp has read tiny[] / before the counter is reused
            if (threadIdx.x == 0) F.ntiny = 0;
            return res;
        }
        if (!(s.lo < s.hi)) return s.lo;
        narrow_level(F, par, hpar, each, s, t);
        exact = true;
    }
}

// the order statistic after `a` (rank r + 1 given a = rank r): a again if more than r + 1 keys are <= a
template <typename Each>
__device__ __forceinline__ float successor(FeatSmemFixed& F, int& par, Each each, float a, int below_outside, int r) {
    int le = 0;
    float ab = INFINITY;
    each([&](float x) { if (x <= a) ++le; else ab = fminf(ab, x); });
    unsigned v[2] = {(unsigned)le, __float_as_uint(ab)};
    block_reduce<RK_IADD, RK_FMIN>(F, par, v);
    return (int)v[0] + below_outside > r + 1 ? a : __uint_as_float(v[1]);
}

// Append-cursor of a thread-private list (slot k of thread t lives at list[k * T + t]).  On the device it is a
// 32-bit shared-window byte address and a push is one predicated STS + IADD (a generic pointer would make the
// compiler rebuild the shared window base on every push).
struct ListCursor {
#ifdef B200LAP_EMUL
    float* p; float* p0; int step;
    __device__ __forceinline__ void init(float* list, int tid, int T, int) { p0 = list + tid; p = p0; step = T; }
    __device__ __forceinline__ void push(float v) { *p = v; p += step; }
    __device__ __forceinline__ bool beyond(int slots) const { return p - p0 > (long)slots * step; }
    __device__ __forceinline__ int count() const { return (int)((p - p0) / step); }
#else
    uint32_t a, a0, step, sh;   // step = 4 T bytes = 1 << sh
    __device__ __forceinline__ void init(float* list, int tid, int T, int log2T) { a0 = smem_u32(list + tid); a = a0; step = 4u * (unsigned)T; sh = 2u + (unsigned)log2T; }
    __device__ __forceinline__ void push(float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); a += step; }
    __device__ __forceinline__ bool beyond(int slots) const { return a - a0 > ((unsigned)slots << sh); }
    __device__ __forceinline__ int count() const { return (int)((a - a0) >> sh); }
#endif
};

// Walk of a thread-private list through its 32-bit shared-window address (LDS + IADD + compare per key).
#ifdef B200LAP_EMUL
template <typename F> __device__ __forceinline__ void walk_list(const ListCursor& c, F f) {
    for (const float* q = c.p0; q != c.p; q += c.step) f(*q);
}
#else
template <typename F> __device__ __forceinline__ void walk_list(const ListCursor& c, F f) {
    for (uint32_t q = c.a0; q != c.a; q += c.step) {
        float x;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(q));
        f(x);
    }
}
#endif

// Upper end of the histogram range of a list admitted by fl(c - L) <= w.  A key a rounding above it lands in the
// last bin (sel_bin256_in clamps), and both list passes use the same bin function, so it need not be exact.
__device__ __forceinline__ float bracket_upper(float L, float w) { return L + w * 1.0001f; }

// ---- warp partial -> CTA cell helpers (lane 0 of every warp publishes with one shared atomic) ----------
__device__ __forceinline__ void warp_min_to(unsigned* cell, float v) {
    const unsigned m = __reduce_min_sync(kFull, f2ord(v));
    if (lane_id() == 0) atomicMin(cell, m);
}
__device__ __forceinline__ void warp_max_to(unsigned* cell, float v) {
    const unsigned m = __reduce_max_sync(kFull, f2ord(v));
    if (lane_id() == 0) atomicMax(cell, m);
}
__device__ __forceinline__ void warp_add_to(int* cell, int v) {
    const int t = (int)__reduce_add_sync(kFull, (unsigned)v);
    if (lane_id() == 0) atomicAdd(cell, t);
}
__device__ __forceinline__ void warp_sum_to(float* slots, float v) {
    v = warp_sum_f(v);
    if (lane_id() == 0) slots[warp_id()] = v;
}
// fixed-order sum of the per-warp partials (any warp)
__device__ __forceinline__ float sum_slots(const float* slots) {
    const int nw = blockDim.x >> 5;
    return warp_sum_f(lane_id() < nw ? slots[lane_id()] : 0.0f);
}

__device__ __forceinline__ float exp2_neg_fast(float t) {   // 2^t for t <= 0, results below the normal range flush to 0
#ifdef B200LAP_EMUL
    return exp2f(t);
#else
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
#endif
}

// Bracket the sample ranks t1 <= t2 (keys samp[0..s), all inside [lo, hi]) by values [L, H].  One histogram
// pass by everybody, the bin search by warp 0 only; an over-full bracket (ties, heavy tails) is narrowed to
// the attained range of its bins and histogrammed again.  L == H signals a bracket of equal keys.
__device__ __forceinline__ void bracket_fast(FeatSmemFixed& F, const float* samp, int s, float lo, float hi, int t1, int t2,
                                             int heavy, float& L, float& H)
{
    const int T = blockDim.x, tid = threadIdx.x, lane = lane_id();
    int before = 0;
    int* h = F.hist[0];
    for (int level = 0;; ++level) {
        if (!(lo < hi)) { L = lo; H = lo; return; }
        const float scale = (float)kSelBins / (hi - lo);
        if (level == 0) {
            for (int i = tid; i < s; i += T) atomicAdd(&h[sel_bin256_in(samp[i], lo, scale)], 1);   // every key is inside [lo, hi]
        } else {
            for (int i = tid; i < s; i += T) {
                const float x = samp[i];
                if (x >= lo && x <= hi) atomicAdd(&h[sel_bin256(x, lo, scale)], 1);
            }
        }
        __syncthreads();
        if (warp_id() == 0) {
            int b1, bb1, c1, b2, bb2, c2;
            find_bin(h, t1 - before, b1, bb1, c1);
            find_bin(h, t2 - before, b2, bb2, c2);
            const int4 z = {0, 0, 0, 0};
            reinterpret_cast<int4*>(h)[2 * lane] = z;
            reinterpret_cast<int4*>(h)[2 * lane + 1] = z;
            if (lane == 0) {
                F.bc_i[0] = b1; F.bc_i[1] = b2; F.bc_i[2] = bb1; F.bc_i[3] = bb2 + c2 - bb1;
                F.cell_h[0] = 0xffffffffu; F.cell_h[1] = 0u;
            }
        }
        __syncthreads();
        const int b1 = F.bc_i[0], b2 = F.bc_i[1], bb1 = F.bc_i[2], inside = F.bc_i[3];
        if (inside <= heavy || level == 2) {
            // bin edges, pushed outwards a little: the exact ranks are resolved later, the edges only have to enclose them
            const float binw = (hi - lo) * (1.0f / (float)kSelBins);
            L = lo + ((float)b1 - 0.02f) * binw;
            H = lo + ((float)b2 + 1.02f) * binw;
            if (!(L < H)) H = nextafterf(L, INFINITY);
            return;
        }
        float mn = INFINITY, mx = -INFINITY;
        for (int i = tid; i < s; i += T) {
            const float x = samp[i];
            if (x >= lo && x <= hi) {
                const int b = sel_bin256(x, lo, scale);
                if (b >= b1 && b <= b2) { mn = fminf(mn, x); mx = fmaxf(mx, x); }
            }
        }
        warp_min_to(&F.cell_h[0], mn);
        warp_max_to(&F.cell_h[1], mx);
        __syncthreads();
        lo = ord2f(F.cell_h[0]); hi = ord2f(F.cell_h[1]);
        before += bb1;
    }
}

template <bool FAST, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_row_features_smem(FeatSmemArgs a)
{
    B200LAP_DYN_SMEM(smem_raw);
    FeatSmemFixed& F = *reinterpret_cast<FeatSmemFixed*>(smem_raw);
    const int T = blockDim.x, tid = threadIdx.x, n = a.n, lane = lane_id();
    const bool leader = warp_id() == 0;
    float* rowbuf = reinterpret_cast<float*>(smem_raw + sizeof(FeatSmemFixed));
    float* list = rowbuf + (size_t)a.nbuf * a.row_floats;
    float* samp = list + (size_t)a.kcap * T;
    const long long total_rows = a.rows_in ? (long long)*a.rows_in_count : (long long)a.batch * n;
    int par = 0, hpar = 0;

    for (int i = tid; i < 2 * kSelBins; i += T) (&F.hist[0][0])[i] = 0;
    auto reset_cells = [&](int pp) {      // by the lanes of one warp
        if (lane < 8) F.cell_u[pp][lane] = (lane == 0 || lane == 2) ? 0xffffffffu : 0u;
        if (lane < 12) F.cell_i[pp][lane] = 0;
    };
    if (leader) { reset_cells(0); reset_cells(1); }
    if (tid == 0) { F.ncand = 0; F.ntiny = 0; }
#ifndef B200LAP_EMUL
    const bool bulk = FAST && a.use_bulk;
    if (bulk && tid == 0) { mbar_init(&F.mbar[0], 1); mbar_init(&F.mbar[1], 1); mbar_fence_init(); }
#else
    const bool bulk = false;
#endif
    __syncthreads();

    auto actual = [&](long long q) -> long long { return a.rows_in ? (long long)a.rows_in[q] : q; };
    auto row_ptr = [&](long long q) { const long long r = actual(q); return a.C + (r / n) * a.inst_stride + (r % n) * (long long)a.ld; };
    // start the arrival of row r in buffer `slot` (bulk: thread 0 only, asynchronous; else cooperative + barrier later)
    auto start_row = [&](long long r, int slot) {
        float* dst = rowbuf + (size_t)slot * a.row_floats;
#ifndef B200LAP_EMUL
        if (bulk) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(&F.mbar[slot], (unsigned)n * 4u);
                bulk_g2s(dst, row_ptr(r), (unsigned)n * 4u, &F.mbar[slot]);
            }
            return;
        }
#endif
        const float* src = row_ptr(r);
        for (int i = tid; i < n; i += T) dst[i] = __ldcs(src + i);
    };
    unsigned phase[2] = {0u, 0u};

    const int n4 = n >> 2;
    const int s = a.nsamp, st = n / s;
    const bool st_pow2 = (st & (st - 1)) == 0;
    const int r1 = (n - 1) >> 1, r2 = n >> 1;   // numpy median: mean of these two order statistics
    int ksel = a.topk > 10 ? a.topk : 10;
    if (ksel > n) ksel = n;
    if (ksel > kTopKMax) ksel = kTopKMax;
    const int nw = T >> 5;
    const int log2T = 31 - __clz(T);
    const int per_warp = (ksel + nw - 1) / nw;
    const double inv_n_d = 1.0 / (double)n;
    const float inv_n = (float)inv_n_d;
    const int sm1 = (int)(((long long)r1 * s) / n), sm2 = (int)(((long long)r2 * s) / n);
    const int st1 = max(0, sm1 - a.delta), st2 = min(s - 1, sm2 + a.delta);
    const int heavy = (st2 - st1) + (st2 - st1) / 2 + s / 32 + 8;   // a bracket holding more sample keys than this is narrowed again
    const int owned = FAST ? 4 * ((n4 - tid + T - 1) / T) : (n - tid + T - 1) / T;   // entries this thread owns
    constexpr float kNegLog2e = -1.4426950408889634f;

    long long r = blockIdx.x;
    if (r < total_rows) start_row(r, 0);
    int it = 0;
    for (; r < total_rows; r += gridDim.x, ++it) {
        const int slot = a.nbuf == 2 ? (it & 1) : 0;
        const int p = it & 1;
        unsigned* CU = F.cell_u[p];
        int* CI = F.cell_i[p];
        const long long rnext = r + gridDim.x;
        if (a.nbuf == 2 && rnext < total_rows) start_row(rnext, slot ^ 1);
#ifndef B200LAP_EMUL
        if (bulk) { mbar_wait(&F.mbar[slot], phase[slot]); phase[slot] ^= 1u; }
        else __syncthreads();
#else
        __syncthreads();
#endif
        const float* buf = rowbuf + (size_t)slot * a.row_floats;
        const long long ra = actual(r);
        const int b = (int)(ra / n), row = (int)(ra % n);
        const float* cm = a.colmin + (size_t)b * n;

        // every key of the row the thread owns, through a key transform (fall-back paths and rare extra passes)
        auto each_row = [&](auto keyf, auto f) {
            if (FAST) {
                for (int g = tid; g < n4; g += T) {
                    const float4 c = reinterpret_cast<const float4*>(buf)[g];
                    f(keyf(c.x)); f(keyf(c.y)); f(keyf(c.z)); f(keyf(c.w));
                }
            } else {
                for (int i = tid; i < n; i += T) f(keyf(buf[i]));
            }
        };
        auto clean_hist = [&]() {   // the fall-back levels leave one histogram dirty
            __syncthreads();
            for (int i = tid; i < 2 * kSelBins; i += T) (&F.hist[0][0])[i] = 0;
            __syncthreads();
        };

        // ---- sample of the row (strided, skewed by one per stride so it is not column periodic)
        {
            float smn = INFINITY, smx = -INFINITY;
            for (int i = tid; i < s; i += T) {
                const float x = buf[i * st + (st_pow2 ? (i & (st - 1)) : (i % st))];
                samp[i] = x;
                smn = fminf(smn, x); smx = fmaxf(smx, x);
            }
            warp_min_to(&CU[0], smn);
            warp_max_to(&CU[1], smx);
        }
        __syncthreads();
        if (leader) reset_cells(p ^ 1);   // the other parity was last read while the previous row finished
        float L, H;
        bracket_fast(F, samp, s, ord2f(CU[0]), ord2f(CU[1]), st1, st2, heavy, L, H);

        // ---- pass 1: min, max, sum, is_col_best, thread minimum; median bracket (count below L, list of [L, H])
        float tmn = INFINITY, tmx = -INFINITY, tsum = 0.0f;
        int notbest = 0, cless = 0, cgt = 0;     // c > colmin;  c < L;  c > L (tie variant)
        ListCursor cur;
        cur.init(list, tid, T, log2T);
        const bool tie = !(L < H);
        unsigned wlim = tie ? 0u : __float_as_uint(H - L) + 1u;
        {
            // counts ride on sign bits (x + (bits >> 31) is one LEA.HI): colmin - c < 0 <=> c is not a column minimum
            auto elem = [&](float c, float m) {
                notbest += __float_as_uint(m - c) >> 31;
                tmn = fminf(tmn, c); tmx = fmaxf(tmx, c);
                tsum += c;
                const unsigned db = __float_as_uint(c - L);
                cless += db >> 31;
                if (db < wlim) cur.push(c);
            };
            auto elem_tie = [&](float c, float m) {
                notbest += __float_as_uint(m - c) >> 31;
                tmn = fminf(tmn, c); tmx = fmaxf(tmx, c);
                tsum += c;
                cless += __float_as_uint(c - L) >> 31;
                cgt += __float_as_uint(L - c) >> 31;
            };
            if (FAST) {
                if (!tie) {
#pragma unroll 2
                    for (int g = tid; g < n4; g += T) {
                        const float4 c = reinterpret_cast<const float4*>(buf)[g];
                        const float4 m = __ldg(reinterpret_cast<const float4*>(cm) + g);
                        if (cur.beyond(a.kcap - 4)) wlim = 0u;
                        elem(c.x, m.x); elem(c.y, m.y); elem(c.z, m.z); elem(c.w, m.w);
                    }
                } else {
#pragma unroll 2
                    for (int g = tid; g < n4; g += T) {
                        const float4 c = reinterpret_cast<const float4*>(buf)[g];
                        const float4 m = __ldg(reinterpret_cast<const float4*>(cm) + g);
                        elem_tie(c.x, m.x); elem_tie(c.y, m.y); elem_tie(c.z, m.z); elem_tie(c.w, m.w);
                    }
                }
            } else {
                for (int i = tid; i < n; i += T) {
                    const float c = buf[i], m = __ldg(cm + i);
                    if (tie) elem_tie(c, m);
                    else { if (cur.beyond(a.kcap - 1)) wlim = 0u; elem(c, m); }
                }
            }
        }
        int mycnt = cur.count();
        {
            // upper bound of the ksel-th smallest entry: per warp the per_warp-th smallest thread minimum
            unsigned key = f2ord(tmn), m = 0;
            for (int q = 0; q < per_warp; ++q) {
                m = __reduce_min_sync(kFull, key);
                const unsigned who = __ballot_sync(kFull, key == m);
                if (lane == __ffs((int)who) - 1) key = 0xffffffffu;
            }
            const unsigned mnw = __reduce_min_sync(kFull, f2ord(tmn)), mxw = __reduce_max_sync(kFull, f2ord(tmx));
            const int nb = (int)__reduce_add_sync(kFull, (unsigned)notbest), cl = (int)__reduce_add_sync(kFull, (unsigned)cless);
            const int in = (int)__reduce_add_sync(kFull, (unsigned)(tie ? owned - cless - cgt : mycnt));
            const bool ov = __any_sync(kFull, !tie && wlim == 0u);
            tsum = warp_sum_f(tsum);
            if (lane == 0) {
                atomicMin(&CU[2], mnw); atomicMax(&CU[3], mxw); atomicMax(&CU[4], m);
                atomicAdd(&CI[0], nb); atomicAdd(&CI[1], cl); atomicAdd(&CI[2], in);
                if (ov) atomicOr(&CI[3], 1);
                F.psum[p][0][warp_id()] = tsum;
            }
        }
        __syncthreads();
        const float mn = ord2f(CU[2]), mx = ord2f(CU[3]), U = ord2f(CU[4]);
        const float sum = sum_slots(F.psum[p][0]);
        const double mean = (double)sum * inv_n_d;
        const float mean_f = (float)mean;

        // ---- median
        auto ident = [](float x) { return x; };
        auto each_row_raw = [&](auto f) { each_row(ident, f); };
        auto each_list = [&](auto f) { walk_list(cur, f); };
        float med_f;
        double med;
        {
            const int below = CI[1], inside = CI[2];
            const bool ok = CI[3] == 0 && r1 >= below && r2 < below + inside;
            float ma = L, mb = L;
            if (ok && !tie) {
                // resolve the two ranks on the list: histogram (all), bin search (warp 0), gather + rank of <= 64 keys (warp 0)
                const float Hb = bracket_upper(L, H - L);
                const float scale = (float)kSelBins / (Hb - L);
                int* h = F.hist[0];
                each_list([&](float x) { atomicAdd(&h[sel_bin256_in(x, L, scale)], 1); });
                __syncthreads();
                if (leader) {
                    int b1, bb1, c1, b2, bb2, c2;
                    find_bin(h, r1 - below, b1, bb1, c1);
                    find_bin(h, r2 - below, b2, bb2, c2);
                    const int4 z = {0, 0, 0, 0};
                    reinterpret_cast<int4*>(h)[2 * lane] = z;
                    reinterpret_cast<int4*>(h)[2 * lane + 1] = z;
                    if (lane == 0) { F.bc_i[4] = b1; F.bc_i[5] = b2; F.bc_i[6] = bb1; F.bc_i[7] = b1 == b2 ? c1 : c1 + c2; }
                }
                __syncthreads();
                const int b1 = F.bc_i[4], b2 = F.bc_i[5], bb1 = F.bc_i[6], cnt = F.bc_i[7];
                if (cnt <= kTinyCap2) {
                    each_list([&](float x) {
                        const int bn = sel_bin256_in(x, L, scale);
                        if (bn == b1 || bn == b2) { const int q = atomicAdd(&F.ntiny, 1); if (q < kTinyCap2) F.tiny[q] = x; }
                    });
                    __syncthreads();
                    if (leader) {
                        const int k1 = r1 - below - bb1, k2 = r2 - below - bb1;
                        const float x0 = lane < cnt ? F.tiny[lane] : INFINITY, x1 = lane + 32 < cnt ? F.tiny[lane + 32] : INFINITY;
                        int rk0 = 0, rk1 = 0;
                        for (int q = 0; q < cnt; ++q) {
                            const float o = F.tiny[q];
                            rk0 += (o < x0) || (o == x0 && q < lane);
                            rk1 += (o < x1) || (o == x1 && q < lane + 32);
                        }
                        if (lane < cnt && rk0 == k1) F.bc_f[0] = x0;
                        if (lane + 32 < cnt && rk1 == k1) F.bc_f[0] = x1;
                        if (lane < cnt && rk0 == k2) F.bc_f[1] = x0;
                        if (lane + 32 < cnt && rk1 == k2) F.bc_f[1] = x1;
                        if (lane == 0) F.ntiny = 0;
                    }
                    __syncthreads();
                    ma = F.bc_f[0]; mb = F.bc_f[1];
                } else {
                    ma = select_rank(F, par, hpar, each_list, SelState{L, Hb, below, inside}, r1, false);
                    mb = r2 == r1 ? ma : successor(F, par, each_list, ma, below, r1);
                    clean_hist();
                }
            } else if (!ok) {
                ma = select_rank(F, par, hpar, each_row_raw, SelState{mn, mx, 0, n}, r1, true);
                mb = r2 == r1 ? ma : successor(F, par, each_row_raw, ma, 0, r1);
                clean_hist();
            }
            med = ((double)ma + (double)mb) * 0.5;
            med_f = (float)med;
        }

        // ---- MAD bracket from the sample deviations (keys |c - med| in binary32: monotone rounding)
        for (int i = tid; i < s; i += T) samp[i] = fabsf(samp[i] - med_f);
        float dhi = fmaxf(mx - med_f, med_f - mn);
        if (!(dhi > 0.0f)) dhi = 0.0f;
        float L2, H2;
        bracket_fast(F, samp, s, 0.0f, dhi, st1, st2, heavy, L2, H2);

        // ---- pass 2: sum z^2 (variance), exp sums, top-k candidates, MAD bracket
        float tzz = 0.0f, tes = 0.0f, tew = 0.0f;
        int cless2 = 0, cgt2 = 0;
        cur.init(list, tid, T, log2T);
        const bool tie2 = !(L2 < H2);
        unsigned wlim2 = tie2 ? 0u : __float_as_uint(H2 - L2) + 1u;
        {
            auto common = [&](float c) {
                const float z = c - mn;
                const float ex = exp2_neg_fast(z * kNegLog2e);
                tzz = fmaf(z, z, tzz);
                tes += ex;
                tew = fmaf(ex, z, tew);
            };
            // the list keeps the signed deviation; readers take the absolute value
            auto elem = [&](float c) {
                common(c);
                const float dv = c - med_f;
                const unsigned db = __float_as_uint(fabsf(dv) - L2);
                cless2 += db >> 31;
                if (db < wlim2) cur.push(dv);
            };
            auto elem_tie = [&](float c) {
                common(c);
                const float key = fabsf(c - med_f);
                cless2 += __float_as_uint(key - L2) >> 31;
                cgt2 += __float_as_uint(L2 - key) >> 31;
            };
            // candidates for the k smallest: one test per group of entries, the rare hit walks the group
            auto cand4 = [&](const float4& c) {
                if (fminf(fminf(c.x, c.y), fminf(c.z, c.w)) < U) {
                    const float v[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (v[q] < U) { const int t = atomicAdd(&F.ncand, 1); if (t < kCandCap) F.cand[t] = v[q]; }
                }
            };
            if (FAST) {
                if (!tie2) {
#pragma unroll 2
                    for (int g = tid; g < n4; g += T) {
                        const float4 c = reinterpret_cast<const float4*>(buf)[g];
                        if (cur.beyond(a.kcap - 4)) wlim2 = 0u;
                        elem(c.x); elem(c.y); elem(c.z); elem(c.w);
                        cand4(c);
                    }
                } else {
#pragma unroll 2
                    for (int g = tid; g < n4; g += T) {
                        const float4 c = reinterpret_cast<const float4*>(buf)[g];
                        elem_tie(c.x); elem_tie(c.y); elem_tie(c.z); elem_tie(c.w);
                        cand4(c);
                    }
                }
            } else {
                for (int i = tid; i < n; i += T) {
                    const float c = buf[i];
                    if (tie2) elem_tie(c);
                    else { if (cur.beyond(a.kcap - 1)) wlim2 = 0u; elem(c); }
                    if (c < U) { const int t = atomicAdd(&F.ncand, 1); if (t < kCandCap) F.cand[t] = c; }
                }
            }
        }
        mycnt = cur.count();
        {
            const int cl = (int)__reduce_add_sync(kFull, (unsigned)cless2);
            const int in = (int)__reduce_add_sync(kFull, (unsigned)(tie2 ? owned - cless2 - cgt2 : mycnt));
            const bool ov = __any_sync(kFull, !tie2 && wlim2 == 0u);
            tzz = warp_sum_f(tzz); tes = warp_sum_f(tes); tew = warp_sum_f(tew);
            if (lane == 0) {
                atomicAdd(&CI[4], cl); atomicAdd(&CI[5], in);
                if (ov) atomicOr(&CI[6], 1);
                F.psum[p][1][warp_id()] = tzz; F.psum[p][2][warp_id()] = tes; F.psum[p][3][warp_id()] = tew;
            }
        }
        __syncthreads();

        // ---- k smallest: when the group-minimum bound left too many candidates, take the exact ksel-th smallest instead
        float thr = U;
        if (F.ncand > kCandCap) {
            __syncthreads();
            if (tid == 0) F.ncand = 0;
            thr = select_rank(F, par, hpar, each_row_raw, SelState{mn, mx, 0, n}, ksel - 1, true);
            clean_hist();
            each_row_raw([&](float c) { if (c < thr) { const int t = atomicAdd(&F.ncand, 1); if (t < kCandCap) F.cand[t] = c; } });
            __syncthreads();
        }
        const int nc = F.ncand;

        // ---- rare extra passes (uniform conditions): near-best count when 1.1*min reaches the candidate bound,
        //      exact variance under cancellation, exp sum without the exact ones for an isolated minimum
        const float near_thr = near_threshold(mn, a.torch_mode);
        const float zz = sum_slots(F.psum[p][1]), es = sum_slots(F.psum[p][2]);
        const double dmean = mean - (double)mn;
        const double var_z = (double)zz * inv_n_d - dmean * dmean;
        const bool need_near = !(near_thr < thr);
        const bool need_var = !(var_z * 20.0 > dmean * dmean);
        const bool need_exp = es < 4.0f;
        if (need_near || need_var || need_exp) {
            int nnear = 0, ones = 0;
            float tss = 0.0f, small = 0.0f;
            each_row_raw([&](float c) {
                nnear += (c <= near_thr);
                const float dl = c - mean_f;
                tss = fmaf(dl, dl, tss);
                const float z = c - mn;
                if (z > 0.0f) small += exp2_neg_fast(z * kNegLog2e); else ++ones;
            });
            warp_add_to(&CI[7], nnear);
            warp_add_to(&CI[8], ones);
            warp_sum_to(F.psum[p][4], tss);
            warp_sum_to(F.psum[p][5], small);
            __syncthreads();
        }

        // ---- MAD: ranks on the list as above; the leader warp finishes the row
        {
            const int below = CI[4], inside = CI[5];
            const bool ok = CI[6] == 0 && r1 >= below && r2 < below + inside;
            float da = L2, db2 = L2;
            bool tiny_pending = false;
            int k1 = 0, k2 = 0, cnt = 0;
            auto each_list_abs = [&](auto f) { walk_list(cur, [&](float x) { f(fabsf(x)); }); };
            if (ok && !tie2) {
                const float Hb = bracket_upper(L2, H2 - L2);
                const float scale = (float)kSelBins / (Hb - L2);
                int* h = F.hist[0];
                each_list_abs([&](float x) { atomicAdd(&h[sel_bin256_in(x, L2, scale)], 1); });
                __syncthreads();
                if (leader) {
                    int b1, bb1, c1, b2, bb2, c2;
                    find_bin(h, r1 - below, b1, bb1, c1);
                    find_bin(h, r2 - below, b2, bb2, c2);
                    const int4 z = {0, 0, 0, 0};
                    reinterpret_cast<int4*>(h)[2 * lane] = z;
                    reinterpret_cast<int4*>(h)[2 * lane + 1] = z;
                    if (lane == 0) { F.bc_i[8] = b1; F.bc_i[9] = b2; F.bc_i[10] = bb1; F.bc_i[11] = b1 == b2 ? c1 : c1 + c2; }
                }
                __syncthreads();
                const int b1 = F.bc_i[8], b2 = F.bc_i[9], bb1 = F.bc_i[10];
                cnt = F.bc_i[11];
                if (cnt <= kTinyCap2) {
                    each_list_abs([&](float x) {
                        const int bn = sel_bin256_in(x, L2, scale);
                        if (bn == b1 || bn == b2) { const int q = atomicAdd(&F.ntiny, 1); if (q < kTinyCap2) F.tiny[q] = x; }
                    });
                    tiny_pending = true;
                    k1 = r1 - below - bb1; k2 = r2 - below - bb1;
                } else {
                    da = select_rank(F, par, hpar, each_list_abs, SelState{L2, Hb, below, inside}, r1, false);
                    db2 = r2 == r1 ? da : successor(F, par, each_list_abs, da, below, r1);
                    clean_hist();
                }
            } else if (!ok) {
                auto devkey = [&](float x) { return fabsf(x - med_f); };
                auto each_row_dev = [&](auto f) { each_row(devkey, f); };
                da = select_rank(F, par, hpar, each_row_dev, SelState{0.0f, dhi, 0, n}, r1, false);
                db2 = r2 == r1 ? da : successor(F, par, each_row_dev, da, 0, r1);
                clean_hist();
            }
            __syncthreads();   // tiny[] / cand[] complete; every pass over the row buffer and the lists is finished

            if (leader) {
                if (tiny_pending) {
                    const float x0 = lane < cnt ? F.tiny[lane] : INFINITY, x1 = lane + 32 < cnt ? F.tiny[lane + 32] : INFINITY;
                    int rk0 = 0, rk1 = 0;
                    for (int q = 0; q < cnt; ++q) {
                        const float o = F.tiny[q];
                        rk0 += (o < x0) || (o == x0 && q < lane);
                        rk1 += (o < x1) || (o == x1 && q < lane + 32);
                    }
                    if (lane < cnt && rk0 == k1) F.bc_f[2] = x0;
                    if (lane + 32 < cnt && rk1 == k1) F.bc_f[2] = x1;
                    if (lane < cnt && rk0 == k2) F.bc_f[3] = x0;
                    if (lane + 32 < cnt && rk1 == k2) F.bc_f[3] = x1;
                    __syncwarp();
                    da = F.bc_f[2]; db2 = F.bc_f[3];
                }
                // the ksel smallest entries, ascending: candidates strictly below thr, then copies of thr
                int nearc = 0;
                for (int t = lane; t < (nc > ksel ? nc : ksel); t += 32) {
                    if (t < nc) {
                        const float mine = F.cand[t];
                        nearc += (mine <= near_thr);
                        int rank = 0;
                        for (int q = 0; q < nc; ++q) {
                            const float o = F.cand[q];
                            rank += (o < mine) || (o == mine && q < t);
                        }
                        if (rank < ksel) F.sorted[rank] = mine;
                    } else {
                        F.sorted[t] = thr;
                    }
                }
                nearc = (int)__reduce_add_sync(kFull, (unsigned)nearc);
                const float ew = sum_slots(F.psum[p][3]);
                const float ss_exact = sum_slots(F.psum[p][4]), small = sum_slots(F.psum[p][5]);
                __syncwarp();
                float* f = a.feat + ((size_t)b * n + row) * kFeatDim;
                if (lane == 0) {
                    F.ncand = 0; F.ntiny = 0;
                    const int near = need_near ? CI[7] : nearc;
                    const int colbest = n - CI[0];
                    double mad = ((double)da + (double)db2) * 0.5;
                    if (mad < 1e-9) mad = 1e-9;
                    float gap = 0.0f, comp = 0.0f, diffi = 0.0f;
                    if (n >= 2) {
                        gap = F.sorted[1] - F.sorted[0];
                        const double range = (double)mx - (double)mn;
                        comp = (float)((double)gap / (range + 1e-9));
                        diffi = (float)(1.0 / (range / (double)(n - 1) + 1e-9));
                    }
                    const int k10 = n < 10 ? n : 10;
                    double km = 0.0;
                    for (int q = 0; q < k10; ++q) km += (double)F.sorted[q];
                    km /= (double)k10;
                    double kv = 0.0;
                    for (int q = 0; q < k10; ++q) { const double t = (double)F.sorted[q] - km; kv += t * t; }
                    float ent;
                    if (need_exp) {
                        const double es_d = (double)CI[8] + (double)small, sp = es_d + 1e-9;
                        ent = (float)((es_d / sp) * log(sp) + (double)ew / sp);
                    } else {
                        const float sp = es + 1e-9f;
                        ent = (es / sp) * logf(sp) + ew / sp;
                    }
                    const double var = need_var ? (double)ss_exact / (double)n : var_z;
                    f[0] = mn; f[1] = mx; f[2] = mean_f;
                    f[3] = (float)sqrt(var > 0.0 ? var : 0.0);
                    f[4] = (float)mad;
                    f[5] = ent;
                    f[6] = gap; f[7] = comp;
                    f[8] = (float)km;
                    f[9] = (float)sqrt(kv / (double)k10);
                    f[10] = diffi;
                    f[11] = (float)near * inv_n;
                    f[12] = (float)colbest * inv_n;
                } else if (lane >= 13 && lane < kFeatDim) {
                    f[lane] = a.posenc[(size_t)row * 8 + (lane - 13)];
                }
                if (a.topv) {
                    const int kout = a.topk < n ? a.topk : n;
                    if (lane < a.topk) a.topv[((size_t)b * n + row) * (size_t)a.topk + lane] = lane < kout && lane < ksel ? F.sorted[lane] : INFINITY;
                }
            }
        }
        __syncthreads();   // the leader is done with the cells, cand[] and sorted[] of this row
        if (a.nbuf == 1) {
            // the single buffer is free again: fetch the next row while the other resident CTAs compute
            if (rnext < total_rows) start_row(rnext, 0);
        }
    }
}

// (sin, cos)(2 pi i f / max(1, n-1)) for f = 1, 2, 4, 8 in binary64, stored binary32 (gnn/features.py:21-31)
__global__ void k_posenc_table(int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double denom = (double)(n - 1 > 1 ? n - 1 : 1);
    for (int q = 0; q < 4; ++q) {
        const double ang = 2.0 * 3.14159265358979323846 * (double)i * (double)(1 << q) / denom;
        out[(size_t)i * 8 + 2 * q] = (float)sin(ang);
        out[(size_t)i * 8 + 2 * q + 1] = (float)cos(ang);
    }
}

}  // namespace b200lap
