// features.cuh -- the 21-D row features and the per-row top-k selection in ONE read of C.
//
// Reference: gnn/features.py:161-243 (compute_row_features), :21-31 (_positional_encodings);
// the top-k values feed gnn/one_gnn.py:143-147 (topk of cost - u_pre: only the VALUES are used,
// and subtracting a per-row constant is monotone, so the k smallest raw entries are selected
// here, before the MLP runs -- SURVEY.md App. C.2).
//
// One CTA of <= 256 threads per row; every thread keeps its EPT (up to 64) entries of the row
// in registers (VEC-interleaved 128-bit loads), so HBM is read once and every later "pass" is
// register-resident.  The kernel is INSTRUCTION bound (ncu, profiles/): what matters is the
// number of instructions per entry and how much per-thread bookkeeping rides on top, hence few
// warps per row, out-of-line reductions, and one code copy of the selection machinery.
//
// Exact order statistics (median, MAD, k-th smallest) without shared-memory histograms: a
// selection level classifies the member keys into 16 value-linear bins (monotone in the value,
// so bins are ordered) by bumping 8-bit fields of two packed 64-bit registers, widens them to
// 16-bit fields, sums them with warp redux instructions (within the warp, then across the
// warps' rows), picks the bin holding the wanted rank, and then either ranks the <= 128
// survivors exhaustively or compacts them (ballot offsets, no atomics) into a shared list that
// the next level works on.  The k smallest entries first use the largest of k group minima as
// an upper bound, which leaves ~90 candidates on random data.
//
// Statistics are accumulated per thread in the storage type and across threads in binary64; exp
// uses the SFU; the entropy uses -sum p log p = (S/S') log S' + sum e (c - z) / S' (one exp per
// entry).  Tolerance 1e-4 relative (+1e-7 absolute), stated in the tests.
// Entries must be finite (a +inf slot marks "past the end of the row").
#pragma once
#include "common.cuh"
#include "frontend.cuh"   // RowLoad, owned_col

namespace b200lap {

constexpr int kFeatDim = 21;
constexpr int kTopKMax = 32;
constexpr int kBins = 16;
constexpr int kListCap = 1024;
constexpr int kSmallCap = 128;
constexpr int kFeatThreads = 256;

struct FeatShared {
    double red[2][4][32];
    unsigned int pk[2][32][kBins / 2];   // per-warp bin counts, two 16-bit fields per word
    double gmin[128];                    // minima of 8-lane groups (top-k bound)
    double list[kListCap];
    double small[kSmallCap];
    double sorted[kTopKMax];
    int nsmall;
    double result;
};

enum { OP_MIN = 0, OP_MAX = 1, OP_SUM = 2 };

template <int OP> __device__ __forceinline__ double red_op(double a, double b) {
    if (OP == OP_MIN) return b < a ? b : a;
    if (OP == OP_MAX) return b > a ? b : a;
    return a + b;
}
template <int OP> __device__ __forceinline__ double red_identity() {
    if (OP == OP_MIN) return INFINITY;
    if (OP == OP_MAX) return -INFINITY;
    return 0.0;
}
template <int OP> __device__ __forceinline__ double warp_red(double v, int first = 16) {
    for (int o = first; o > 0; o >>= 1) v = red_op<OP>(v, shfl_xor_d(v, o));
    return v;
}

// four independent block reductions behind one barrier (out of line: called ~12 times per row)
struct Quad { double a, b, c, d; };
template <int O0, int O1, int O2, int O3>
__device__ __noinline__ Quad block_red4_impl(FeatShared& S, int par, Quad q) {
    q.a = warp_red<O0>(q.a); q.b = warp_red<O1>(q.b); q.c = warp_red<O2>(q.c); q.d = warp_red<O3>(q.d);
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) {
        S.red[par][0][warp_id()] = q.a; S.red[par][1][warp_id()] = q.b;
        S.red[par][2][warp_id()] = q.c; S.red[par][3][warp_id()] = q.d;
    }
    __syncthreads();
    int first = 1;
    while (first < nw) first <<= 1;
    first >>= 1;                          // butterfly over the next power of two >= nw lanes only
    const bool in = lane_id() < nw;
    q.a = warp_red<O0>(in ? S.red[par][0][lane_id()] : red_identity<O0>(), first);
    q.b = warp_red<O1>(in ? S.red[par][1][lane_id()] : red_identity<O1>(), first);
    q.c = warp_red<O2>(in ? S.red[par][2][lane_id()] : red_identity<O2>(), first);
    q.d = warp_red<O3>(in ? S.red[par][3][lane_id()] : red_identity<O3>(), first);
    q.a = __shfl_sync(kFull, q.a, 0); q.b = __shfl_sync(kFull, q.b, 0);   // lanes beyond the butterfly hold partials
    q.c = __shfl_sync(kFull, q.c, 0); q.d = __shfl_sync(kFull, q.d, 0);
    return q;
}
template <int O0, int O1, int O2, int O3>
__device__ __forceinline__ void block_red4(FeatShared& S, int& par, double& a, double& b, double& c, double& d) {
    par ^= 1;
    const Quad q = block_red4_impl<O0, O1, O2, O3>(S, par, Quad{a, b, c, d});
    a = q.a; b = q.b; c = q.c; d = q.d;
}

template <typename KT> __device__ __forceinline__ int sel_bin(KT x, KT lo, KT scale) {
    const KT t = (x - lo) * scale;
    return t >= (KT)(kBins - 1) ? kBins - 1 : (t > (KT)0 ? (int)t : 0);
}

// ---- one selection level: bin counts of the member keys -> (bin, keys before it, keys in it) ------
struct BinPick { int bin, before, count, wbase; };

__device__ __forceinline__ unsigned int pick_word(const unsigned int (&w)[kBins / 2], int idx) {
    unsigned int v = w[0];
#pragma unroll
    for (int q = 1; q < kBins / 2; ++q) v = idx == q ? w[q] : v;
    return v;
}

// c0/c1: the thread's counts as 8-bit fields (bins 0-7, 8-15); a thread owns <= 64 member keys
__device__ __noinline__ BinPick pick_bin_finish(unsigned long long c0, unsigned long long c1, int r, FeatShared& S, int ppar)
{
    unsigned int w[kBins / 2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        w[q] = (unsigned int)((c0 >> (16 * q)) & 0xffull) | ((unsigned int)((c0 >> (16 * q + 8)) & 0xffull) << 16);
        w[4 + q] = (unsigned int)((c1 >> (16 * q)) & 0xffull) | ((unsigned int)((c1 >> (16 * q + 8)) & 0xffull) << 16);
    }
#pragma unroll
    for (int q = 0; q < kBins / 2; ++q) w[q] = __reduce_add_sync(kFull, w[q]);   // <= 32*64 per 16-bit field
    const int l = lane_id();
    if (l < kBins / 2) S.pk[ppar][warp_id()][l] = pick_word(w, l);
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    // column sums over the warps: lane l holds warp l's row, one redux per packed word (<= 32768 per field)
    unsigned int t[kBins / 2];
#pragma unroll
    for (int q = 0; q < kBins / 2; ++q) t[q] = __reduce_add_sync(kFull, l < nw ? S.pk[ppar][l][q] : 0u);
    const int total = l < kBins ? (int)((pick_word(t, l >> 1) >> (16 * (l & 1))) & 0xffffu) : 0;
    int incl = total;
#pragma unroll
    for (int o = 1; o < kBins; o <<= 1) {
        const int u = __shfl_up_sync(kFull, incl, o);
        if (l >= o) incl += u;
    }
    const int start = incl - total;
    const unsigned int hit = __ballot_sync(kFull, l < kBins && total > 0 && r >= start && r < start + total);
    BinPick p;
    p.bin = hit ? __ffs((int)hit) - 1 : 0;
    p.before = __shfl_sync(kFull, start, p.bin);
    p.count = hit ? __shfl_sync(kFull, total, p.bin) : 0;
    // keys of that bin held by the warps before this one (compaction offset)
    const int mine = l < warp_id() ? (int)((S.pk[ppar][l][p.bin >> 1] >> (16 * (p.bin & 1))) & 0xffffu) : 0;
    p.wbase = __reduce_add_sync(kFull, mine);
    return p;
}

// exhaustive rank among <= kSmallCap collected keys; every thread returns the key of rank `want`
__device__ __noinline__ double small_rank_select(FeatShared& S, int count, int want)
{
    const int T = blockDim.x, tid = threadIdx.x;
    __syncthreads();   // S.small complete
    for (int t = tid; t < count; t += T) {
        const double mine = S.small[t];
        int rank = 0;
        for (int q = 0; q < count; ++q) {
            const double o = S.small[q];
            rank += (o < mine) || (o == mine && q < t);
        }
        if (rank == want) S.result = mine;
    }
    __syncthreads();
    const double res = S.result;
    return res;
}

// Exact r-th smallest (0-based) of the keys the thread holds in registers (kf(e), e < EPT; slots
// past the row end hold +inf and never pass the range test).  All real keys lie in [lo, hi].
// Uniform control flow: every thread takes the same path.
template <typename KT, int EPT, typename KeyF>
__device__ __forceinline__ KT block_select(KeyF kf, int r, KT lo, KT hi, FeatShared& S, int& par, int& ppar)
{
    const int T = blockDim.x, tid = threadIdx.x;
    int M = 0;   // > 0: the members are S.list[0..M)
    while (true) {
        if (!(lo < hi)) return lo;
        if (tid == 0) S.nsmall = 0;
        const KT scale = (KT)kBins / (hi - lo);
        unsigned long long c0 = 0ull, c1 = 0ull;
        auto count_key = [&](KT k) {
            if (k >= lo && k <= hi) {
                const int b = sel_bin(k, lo, scale);
                const unsigned long long one = 1ull << (8 * (b & 7));
                if (b < 8) c0 += one; else c1 += one;
            }
        };
        if (M) {
            for (int idx = tid; idx < M; idx += T) count_key((KT)S.list[idx]);
        } else {
#pragma unroll
            for (int e = 0; e < EPT; ++e) count_key(kf(e));
        }
        ppar ^= 1;
        const BinPick p = pick_bin_finish(c0, c1, r, S, ppar);
        if (p.count == 0) return lo;   // unreachable for consistent inputs (NaN keys)
        r -= p.before;
        auto member = [&](KT k) { return k >= lo && k <= hi && sel_bin(k, lo, scale) == p.bin; };
        if (p.count <= kSmallCap) {
            auto grab = [&](KT k) { if (member(k)) S.small[atomicAdd(&S.nsmall, 1)] = (double)k; };
            if (M) {
                for (int idx = tid; idx < M; idx += T) grab((KT)S.list[idx]);
            } else {
#pragma unroll
                for (int e = 0; e < EPT; ++e) grab(kf(e));
            }
            return (KT)small_rank_select(S, p.count, r);
        }
        KT bmin = (KT)INFINITY, bmax = (KT)(-INFINITY);
        if (!M && p.count <= kListCap) {
            // compact the bin into the shared list (ballot offsets), then work on the list
            int run = p.wbase;
            const unsigned int lt = (1u << lane_id()) - 1u;
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const KT k = kf(e);
                const bool take = member(k);
                const unsigned int m = __ballot_sync(kFull, take);
                if (take) {
                    S.list[run + __popc(m & lt)] = (double)k;
                    bmin = k < bmin ? k : bmin;
                    bmax = k > bmax ? k : bmax;
                }
                run += __popc(m);
            }
            M = p.count;
        } else {
            // too many keys in the bin: shrink the range to the bin's exact extent and go again
            auto ext = [&](KT k) {
                if (member(k)) {
                    bmin = k < bmin ? k : bmin;
                    bmax = k > bmax ? k : bmax;
                }
            };
            if (M) {
                for (int idx = tid; idx < M; idx += T) ext((KT)S.list[idx]);
            } else {
#pragma unroll
                for (int e = 0; e < EPT; ++e) ext(kf(e));
            }
        }
        double dmin = (double)bmin, dmax = (double)bmax, z0 = 0.0, z1 = 0.0;
        block_red4<OP_MIN, OP_MAX, OP_SUM, OP_SUM>(S, par, dmin, dmax, z0, z1);   // the barrier also publishes the list
        lo = (KT)dmin;
        hi = (KT)dmax;
    }
}

template <typename CT> __device__ __forceinline__ CT round_down_to(double x);
template <> __device__ __forceinline__ float round_down_to<float>(double x) {
#ifdef B200LAP_EMUL
    float f = (float)x;
    return (double)f > x ? nextafterf(f, -INFINITY) : f;
#else
    return __double2float_rd(x);
#endif
}
template <> __device__ __forceinline__ double round_down_to<double>(double x) { return x; }

template <typename CT> __device__ __forceinline__ CT fast_exp_neg(CT x);   // exp(-x), x >= 0
template <> __device__ __forceinline__ float fast_exp_neg<float>(float x) { return __expf(-x); }
template <> __device__ __forceinline__ double fast_exp_neg<double>(double x) { return exp(-x); }

__device__ __noinline__ void write_row_features(float* f, int n, int row, double mn, double mx, double mean, double ssq,
                                                double mad, double esum, double ewsum, double near, double cb,
                                                const double* sorted)
{
    double gap = 0.0, comp = 0.0, diffi = 0.0;
    if (n >= 2) {
        gap = sorted[1] - sorted[0];
        comp = gap / ((mx - mn) + 1e-9);
        diffi = 1.0 / ((mx - mn) / (double)(n - 1) + 1e-9);
    }
    const int k10 = n < 10 ? n : 10;
    double km = 0.0;
    for (int q = 0; q < k10; ++q) km += sorted[q];
    km /= (double)k10;
    double kv = 0.0;
    for (int q = 0; q < k10; ++q) { const double t = sorted[q] - km; kv += t * t; }
    kv = sqrt(kv / (double)k10);
    const double sp = esum + 1e-9;
    const double ent = (esum / sp) * log(sp) + ewsum / sp;
    f[0] = (float)mn;
    f[1] = (float)mx;
    f[2] = (float)mean;
    f[3] = (float)sqrt(ssq / (double)n);
    f[4] = (float)mad;
    f[5] = (float)ent;
    f[6] = (float)gap;
    f[7] = (float)comp;
    f[8] = (float)km;
    f[9] = (float)kv;
    f[10] = (float)diffi;
    f[11] = (float)(near / (double)n);
    f[12] = (float)(cb / (double)n);
    const double denom = (double)(n - 1 > 1 ? n - 1 : 1);
    for (int q = 0; q < 4; ++q) {
        const double fr = (double)(1 << q);
        const double ang = 2.0 * 3.14159265358979323846 * (double)row * fr / denom;
        f[13 + 2 * q] = (float)sin(ang);
        f[14 + 2 * q] = (float)cos(ang);
    }
}

template <typename CT, int VEC, int EPT, int MAXT>
__global__ void __launch_bounds__(MAXT) k_row_features(
    const CT* __restrict__ C, long long inst_stride, int ld, int n, int topk,
    const CT* __restrict__ colmin /* [B][n] */, float* __restrict__ feat /* [B][n][21] */,
    float* __restrict__ topv /* [B][n][topk] or null */)
{
    __shared__ FeatShared S;
    const int b = blockIdx.y, row = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
    const int nw = T >> 5;
    const CT* crow = C + (size_t)b * inst_stride + (size_t)row * ld;
    const CT* cm = colmin + (size_t)b * n;
    int par = 0, ppar = 0;
    CT cv[EPT];
    int colbest = 0;
    CT tmn = (CT)INFINITY, tmx = (CT)(-INFINITY), tsum = (CT)0;
#pragma unroll
    for (int g = 0; g < EPT / VEC; ++g) {
        const int col = owned_col<VEC>(g * VEC, T, tid);
        if (col < n) {
            RowLoad<CT, VEC>::ld(crow + col, &cv[g * VEC]);
            CT mm[VEC];
            RowLoad<CT, VEC>::ld(cm + col, mm);
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                const CT c = cv[g * VEC + q];
                colbest += (c == mm[q]);
                tmn = c < tmn ? c : tmn;
                tmx = c > tmx ? c : tmx;
                tsum += c;
            }
        } else {
#pragma unroll
            for (int q = 0; q < VEC; ++q) cv[g * VEC + q] = (CT)INFINITY;   // "past the end" marker
        }
    }
    // minima of 8-lane groups feed the top-k bound; published by the first reduction's barrier
    {
        double gm = (double)tmn;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { const double t = shfl_xor_d(gm, o); gm = t < gm ? t : gm; }
        if ((lane_id() & 7) == 0) S.gmin[tid >> 3] = gm;
    }
    double mn = (double)tmn, mx = (double)tmx, sum = (double)tsum, cb = (double)colbest;
    block_red4<OP_MIN, OP_MAX, OP_SUM, OP_SUM>(S, par, mn, mx, sum, cb);
    const double mean = sum / (double)n;
    const CT mean_c = (CT)mean, mn_c = (CT)mn;
    const CT near_thr = round_down_to<CT>(mn * 1.1);   // c <= mn*1.1 in binary64  <=>  c <= round_down(mn*1.1)
    CT tssq = (CT)0, tes = (CT)0, tew = (CT)0;
    int tnear = 0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const CT c = cv[e];
        if (c < (CT)INFINITY) {
            const CT dlt = c - mean_c;
            tssq += dlt * dlt;
            const CT z = c - mn_c;
            const CT ex = fast_exp_neg<CT>(z);
            tes += ex;
            tew += ex * z;
            tnear += (c <= near_thr);
        }
    }
    double ssq = (double)tssq, esum = (double)tes, ewsum = (double)tew, near = (double)tnear;
    block_red4<OP_SUM, OP_SUM, OP_SUM, OP_SUM>(S, par, ssq, esum, ewsum, near);

    // One code copy of the selection machinery serves three jobs (instruction footprint):
    //   job 0: the ksel-th smallest entry (only when the group-minimum bound leaves too many candidates)
    //   job 1: the median of the row            job 2: the median of |c - median|
    // |c - med| is formed in the storage type: rounding is monotone up to one storage ulp of the
    // largest deviation, far inside the feature tolerance.
    CT med_c = (CT)0;
    int mode = 0;   // 0: key = c, 1: key = |c - med_c|
    auto key_of = [&](int e) {
        const CT t = cv[e] - med_c;
        return mode ? (t < (CT)0 ? -t : t) : cv[e];
    };

    // ---- k smallest, ascending (k = what the features (10) and the model (topk) need)
    int ksel = topk > 10 ? topk : 10;
    if (ksel > n) ksel = n;
    if (ksel > kTopKMax) ksel = kTopKMax;
    // upper bound U: the largest of the first ksel group minima (at least ksel entries are <= U)
    double U = INFINITY;
    if (4 * nw >= ksel) {
        double g = lane_id() < ksel ? S.gmin[lane_id()] : -INFINITY;
        g = warp_red<OP_MAX>(g);
        U = g;
    }
    bool need_thr = true;
    {
        int lt = kSmallCap + 1;
        if (U < INFINITY) {
            lt = 0;
#pragma unroll
            for (int e = 0; e < EPT; ++e) lt += ((double)cv[e] < U);
        }
        if (tid == 0) S.nsmall = 0;
        double clt = (double)lt, z0 = 0.0, z1 = 0.0, z2 = 0.0;
        block_red4<OP_SUM, OP_SUM, OP_SUM, OP_SUM>(S, par, clt, z0, z1, z2);
        need_thr = !((int)clt <= kSmallCap && U < INFINITY);
    }
    double med = 0.0, mad = 0.0;
    CT dhi = (CT)0;
#pragma unroll 1
    for (int job = 0; job < 3; ++job) {
        int r = (n - 1) / 2;
        CT lo = (CT)mn, hi = (CT)mx;
        mode = 0;
        if (job == 0) r = ksel - 1;
        if (job == 2) { mode = 1; lo = (CT)0; hi = dhi; }
        CT a;
        if (job == 0 && !need_thr) a = (CT)U;
        else a = block_select<CT, EPT>(key_of, r, lo, hi, S, par, ppar);
        if (job == 0) {
            const CT thr = a;
            if (need_thr) {
                if (tid == 0) S.nsmall = 0;
                __syncthreads();
            }
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if (cv[e] < thr) S.small[atomicAdd(&S.nsmall, 1)] = (double)cv[e];
            __syncthreads();
            const int c = S.nsmall;   // <= kSmallCap (bound path) or < ksel (select path)
            for (int t = tid; t < (c > ksel ? c : ksel); t += T) {
                if (t < c) {
                    const double mine = S.small[t];
                    int rank = 0;
                    for (int q = 0; q < c; ++q) {
                        const double o = S.small[q];
                        rank += (o < mine) || (o == mine && q < t);
                    }
                    if (rank < ksel) S.sorted[rank] = mine;
                } else {
                    S.sorted[t] = (double)thr;
                }
            }
            __syncthreads();
            if (topv) {
                const int kout = topk < n ? topk : n;
                float* tv = topv + ((size_t)b * n + row) * (size_t)topk;
                for (int t = tid; t < topk; t += T) tv[t] = t < kout && t < ksel ? (float)S.sorted[t] : INFINITY;
            }
            continue;
        }
        // median = mean of the order statistics (n-1)/2 and n/2 (numpy's definition)
        double value = (double)a;
        if (!(n & 1)) {
            int cnt = 0;
            CT ab = (CT)INFINITY;
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const CT ke = key_of(e);
                if (ke <= a) ++cnt;
                else ab = ke < ab ? ke : ab;
            }
            double le = (double)cnt, above = (double)ab, z0 = 0.0, z1 = 0.0;
            block_red4<OP_SUM, OP_MIN, OP_SUM, OP_SUM>(S, par, le, above, z0, z1);
            const double b2 = ((int)le > r + 1) ? (double)a : above;
            value = ((double)a + b2) / 2.0;
        }
        if (job == 1) {
            med = value;
            med_c = (CT)med;
            const CT a1 = (CT)mx - med_c, a2 = med_c - (CT)mn;
            dhi = a1 > a2 ? a1 : a2;
            if (dhi < (CT)0) dhi = (CT)0;
        } else {
            mad = value;
        }
    }
    if (mad < 1e-9) mad = 1e-9;
    if (tid == 0)
        write_row_features(feat + ((size_t)b * n + row) * kFeatDim, n, row, mn, mx, mean, ssq, mad, esum, ewsum, near, cb, S.sorted);
}


// ---- compute_row_features_torch (gnn/features.py:246-351) on top of the NumPy-definition features ---------------
// The torch variant differs from compute_row_features in four places: row_std and k_std are unbiased (torch.std),
// is_col_best counts only the FIRST row attaining each column minimum (argmin + bincount, :319-320) and the near-best
// threshold is multiplied in binary32 (handled inside the feature kernels).
__global__ void k_count_col_argmin(const int* __restrict__ colarg, int n, int* __restrict__ cnt /* [B][n], zeroed */)
{
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) atomicAdd(&cnt[(size_t)b * n + colarg[(size_t)b * n + j]], 1);
}
__global__ void k_feat_torch_fixup(float* __restrict__ feat, const int* __restrict__ cnt, int n, long long rows)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float* f = feat + r * kFeatDim;
    const int k = n < 10 ? n : 10;
    f[3] = n > 1 ? f[3] * sqrtf((float)n / (float)(n - 1)) : NAN;           // torch.std of one element is NaN
    f[9] = k > 1 ? f[9] * sqrtf((float)k / (float)(k - 1)) : NAN;
    f[12] = (float)cnt[r] / (float)n;
}

}  // namespace b200lap
