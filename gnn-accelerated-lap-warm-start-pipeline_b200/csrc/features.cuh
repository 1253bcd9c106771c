// features.cuh -- the 21-D row features and the per-row top-k selection in ONE read of C.
//
// Reference: gnn/features.py:161-243 (compute_row_features), :21-31 (_positional_encodings);
// the top-k values feed gnn/one_gnn.py:143-147 (topk of cost - u_pre: only the VALUES are used,
// and subtracting a per-row constant is monotone, so the k smallest raw entries are selected
// here, before the MLP runs -- SURVEY.md App. C.2).
//
// One CTA per row; every thread keeps its EPT entries of the row in registers (VEC-interleaved,
// 128-bit loads), so HBM is read once and every later "pass" is register-resident.
//
// Exact order statistics (median, MAD, k-th smallest) WITHOUT shared-memory histograms -- a
// shared atomic costs ~2 cycles per element per SM on this part, 25x the cost of counting in
// registers (measured: the first version of this kernel spent 140k cycles per 16384-entry row in
// ATOMS).  A selection level classifies the member keys into 16 value-linear bins (monotone in
// the value, so bins are ordered) by bumping 8-bit fields of two packed 64-bit registers, widens
// them to 16-bit fields, sums them with 8 warp redux instructions and one shared-memory round,
// picks the bin holding the wanted rank, and then either ranks the <= 64 survivors exhaustively,
// or compacts them (ballot offsets, no atomics) into a shared list that the next level works on.
// The k smallest entries use the k-th smallest WARP MINIMUM as an upper bound first, which
// leaves ~20 candidates on random data.
//
// Statistics are accumulated per thread in the storage type and across threads in binary64; exp
// uses the SFU; the entropy uses -sum p log p = (S/S') log S' + sum e (c - z) / S' (one exp per
// entry).  Tolerance 1e-4 relative (+1e-7 absolute), stated in the tests.
#pragma once
#include "common.cuh"
#include "frontend.cuh"   // RowLoad, owned_col

namespace b200lap {

constexpr int kFeatDim = 21;
constexpr int kTopKMax = 32;
constexpr int kBins = 16;
constexpr int kListCap = 1024;
constexpr int kSmallCap = 64;

struct FeatShared {
    double red[2][4][32];
    unsigned int pk[2][32][kBins / 2];   // per-warp bin counts, two 16-bit fields per word
    double wmin[32];
    double list[kListCap];
    double small[kSmallCap];
    double sorted[kTopKMax];
    int wcnt[2][32];
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
template <int OP> __device__ __forceinline__ double warp_red(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_op<OP>(v, shfl_xor_d(v, o));
    return v;
}

// four independent block reductions behind one barrier
template <int O0, int O1, int O2, int O3>
__device__ __forceinline__ void block_red4(FeatShared& S, int& par, double& a, double& b, double& c, double& d) {
    a = warp_red<O0>(a); b = warp_red<O1>(b); c = warp_red<O2>(c); d = warp_red<O3>(d);
    par ^= 1;
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) {
        S.red[par][0][warp_id()] = a; S.red[par][1][warp_id()] = b;
        S.red[par][2][warp_id()] = c; S.red[par][3][warp_id()] = d;
    }
    __syncthreads();
    const bool in = lane_id() < nw;
    a = warp_red<O0>(in ? S.red[par][0][lane_id()] : red_identity<O0>());
    b = warp_red<O1>(in ? S.red[par][1][lane_id()] : red_identity<O1>());
    c = warp_red<O2>(in ? S.red[par][2][lane_id()] : red_identity<O2>());
    d = warp_red<O3>(in ? S.red[par][3][lane_id()] : red_identity<O3>());
}

template <typename KT> __device__ __forceinline__ int sel_bin(KT x, KT lo, KT scale) {
    const KT t = (x - lo) * scale;
    return t >= (KT)(kBins - 1) ? kBins - 1 : (t > (KT)0 ? (int)t : 0);
}

// ---- one selection level: bin counts of the member keys -> (bin, keys before it, keys in it) ------
struct BinPick { int bin, before, count, wbase; };

template <typename KT, typename Each>
__device__ __forceinline__ BinPick pick_bin(Each&& each, KT lo, KT hi, KT scale, int r, FeatShared& S, int& ppar)
{
    unsigned long long c0 = 0ull, c1 = 0ull;   // 8-bit fields: bins 0-7, 8-15 (a thread owns <= 32 keys)
    each([&](KT k) {
        if (k >= lo && k <= hi) {
            const int b = sel_bin(k, lo, scale);
            if (b < 8) c0 += 1ull << (8 * b); else c1 += 1ull << (8 * (b - 8));
        }
    });
    ppar ^= 1;
    unsigned int w[kBins / 2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        w[q] = (unsigned int)((c0 >> (16 * q)) & 0xffull) | ((unsigned int)((c0 >> (16 * q + 8)) & 0xffull) << 16);
        w[4 + q] = (unsigned int)((c1 >> (16 * q)) & 0xffull) | ((unsigned int)((c1 >> (16 * q + 8)) & 0xffull) << 16);
    }
#pragma unroll
    for (int q = 0; q < kBins / 2; ++q) w[q] = __reduce_add_sync(kFull, w[q]);   // <= 32*32 per field
    if (lane_id() < kBins / 2) {
        unsigned int mine = w[0];
#pragma unroll
        for (int q = 1; q < kBins / 2; ++q) mine = lane_id() == q ? w[q] : mine;
        S.pk[ppar][warp_id()][lane_id()] = mine;
    }
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    const int l = lane_id();
    int total = 0;
    if (l < kBins)
        for (int wv = 0; wv < nw; ++wv) total += (int)((S.pk[ppar][wv][l >> 1] >> (16 * (l & 1))) & 0xffffu);
    int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (l >= o) incl += t;
    }
    const int start = incl - total;
    const unsigned int hit = __ballot_sync(kFull, l < kBins && total > 0 && r >= start && r < start + total);
    BinPick p;
    p.bin = hit ? __ffs((int)hit) - 1 : 0;
    p.before = __shfl_sync(kFull, start, p.bin);
    p.count = hit ? __shfl_sync(kFull, total, p.bin) : 0;
    // keys of that bin held by the warps before this one (compaction offset)
    const int mine = l < warp_id() ? (int)((S.pk[ppar][l][p.bin >> 1] >> (16 * (p.bin & 1))) & 0xffffu) : 0;
    p.wbase = warp_sum_i(mine);
    return p;
}

// exhaustive rank among <= kSmallCap collected keys; every thread returns the key of rank `want`
__device__ __forceinline__ double small_rank_select(FeatShared& S, int count, int want)
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

// Exact r-th smallest (0-based) of the keys the thread holds in registers (kf(e), e < EPT, valid
// columns only).  All keys lie in [lo, hi].  Uniform control flow: every thread takes the same path.
template <typename KT, int VEC, int EPT, typename KeyF>
__device__ KT block_select(KeyF kf, int n, int r, KT lo, KT hi, FeatShared& S, int& par, int& ppar)
{
    const int T = blockDim.x, tid = threadIdx.x;
    auto each_reg = [&](auto&& f) {
#pragma unroll
        for (int e = 0; e < EPT; ++e)
            if (owned_col<VEC>(e, T, tid) < n) f(kf(e));
    };
    int M = 0;   // > 0: the members are S.list[0..M)
    auto each_list = [&](auto&& f) {
        for (int idx = tid; idx < M; idx += T) f((KT)S.list[idx]);
    };
    while (true) {
        if (!(lo < hi)) return lo;
        if (tid == 0) S.nsmall = 0;
        const KT scale = (KT)kBins / (hi - lo);
        const BinPick p = M ? pick_bin<KT>(each_list, lo, hi, scale, r, S, ppar) : pick_bin<KT>(each_reg, lo, hi, scale, r, S, ppar);
        if (p.count == 0) return lo;   // unreachable for consistent inputs (NaN keys)
        r -= p.before;
        auto member = [&](KT k) { return k >= lo && k <= hi && sel_bin(k, lo, scale) == p.bin; };
        if (p.count <= kSmallCap) {
            auto grab = [&](KT k) { if (member(k)) S.small[atomicAdd(&S.nsmall, 1)] = (double)k; };
            if (M) each_list(grab); else each_reg(grab);
            return (KT)small_rank_select(S, p.count, r);
        }
        double bmin = INFINITY, bmax = -INFINITY, z0 = 0.0, z1 = 0.0;
        if (!M && p.count <= kListCap) {
            // compact the bin into the shared list (ballot offsets), then work on the list
            int run = p.wbase;
            const unsigned int lt = (1u << lane_id()) - 1u;
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const bool valid = owned_col<VEC>(e, T, tid) < n;
                const KT k = kf(e);
                const bool take = valid && member(k);
                const unsigned int m = __ballot_sync(kFull, take);
                if (take) {
                    S.list[run + __popc(m & lt)] = (double)k;
                    bmin = (double)k < bmin ? (double)k : bmin;
                    bmax = (double)k > bmax ? (double)k : bmax;
                }
                run += __popc(m);
            }
            block_red4<OP_MIN, OP_MAX, OP_SUM, OP_SUM>(S, par, bmin, bmax, z0, z1);   // barrier publishes the list
            M = p.count;
        } else {
            // too many keys in the bin: shrink the range to the bin's exact extent and go again
            auto ext = [&](KT k) {
                if (member(k)) {
                    bmin = (double)k < bmin ? (double)k : bmin;
                    bmax = (double)k > bmax ? (double)k : bmax;
                }
            };
            if (M) each_list(ext); else each_reg(ext);
            block_red4<OP_MIN, OP_MAX, OP_SUM, OP_SUM>(S, par, bmin, bmax, z0, z1);
        }
        lo = (KT)bmin;
        hi = (KT)bmax;
    }
}

// median = mean of the order statistics (n-1)/2 and n/2 (numpy's definition)
template <typename KT, int VEC, int EPT, typename KeyF>
__device__ double block_median(KeyF kf, int n, KT lo, KT hi, FeatShared& S, int& par, int& ppar)
{
    const int T = blockDim.x, tid = threadIdx.x;
    const int r1 = (n - 1) / 2;
    const KT a = block_select<KT, VEC, EPT>(kf, n, r1, lo, hi, S, par, ppar);
    if (n & 1) return (double)a;
    double le = 0.0, above = INFINITY, z0 = 0.0, z1 = 0.0;
    int cnt = 0;
    KT ab = (KT)INFINITY;
#pragma unroll
    for (int e = 0; e < EPT; ++e)
        if (owned_col<VEC>(e, T, tid) < n) {
            const KT ke = kf(e);
            if (ke <= a) ++cnt;
            else ab = ke < ab ? ke : ab;
        }
    le = (double)cnt;
    above = (double)ab;
    block_red4<OP_SUM, OP_MIN, OP_SUM, OP_SUM>(S, par, le, above, z0, z1);
    const double b = ((int)le > r1 + 1) ? (double)a : above;
    return ((double)a + b) / 2.0;
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

template <typename CT, int VEC, int EPT>
__global__ void __launch_bounds__(1024) k_row_features(
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
            for (int q = 0; q < VEC; ++q) cv[g * VEC + q] = (CT)0;
        }
    }
    // warp minima feed the top-k bound; publish them with the first reduction's barrier
    {
        const double wm = warp_min_d((double)tmn);
        if (lane_id() == 0) S.wmin[warp_id()] = wm;
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
        if (owned_col<VEC>(e, T, tid) < n) {
            const CT c = cv[e];
            const CT dlt = c - mean_c;
            tssq += dlt * dlt;
            const CT z = c - mn_c;
            const CT ex = fast_exp_neg<CT>(z);
            tes += ex;
            tew += ex * (z < (CT)3.0e38 ? z : (CT)0);   // e == 0 there; avoids 0 * inf
            tnear += (c <= near_thr);
        }
    }
    double ssq = (double)tssq, esum = (double)tes, ewsum = (double)tew, near = (double)tnear;
    block_red4<OP_SUM, OP_SUM, OP_SUM, OP_SUM>(S, par, ssq, esum, ewsum, near);

    auto raw_key = [&](int e) { return cv[e]; };

    // ---- k smallest, ascending (k = what the features (10) and the model (topk) need)
    int ksel = topk > 10 ? topk : 10;
    if (ksel > n) ksel = n;
    if (ksel > kTopKMax) ksel = kTopKMax;
    {
        // upper bound U: the ksel-th smallest warp minimum (at least ksel entries are <= U)
        double U = INFINITY;
        if (nw >= ksel) {
            const double w = lane_id() < nw ? S.wmin[lane_id()] : INFINITY;
            int rank = 0;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const double o = __shfl_sync(kFull, w, q);
                rank += (o < w) || (o == w && q < lane_id());
            }
            const unsigned int who = __ballot_sync(kFull, rank == ksel - 1);
            U = __shfl_sync(kFull, w, __ffs((int)who) - 1);
        }
        int lt = 0;
        if (U < INFINITY) {
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if (owned_col<VEC>(e, T, tid) < n) lt += ((double)cv[e] < U);
        } else {
            lt = kSmallCap + 1;
        }
        if (tid == 0) S.nsmall = 0;
        double clt = (double)lt, z0 = 0.0, z1 = 0.0, z2 = 0.0;
        block_red4<OP_SUM, OP_SUM, OP_SUM, OP_SUM>(S, par, clt, z0, z1, z2);
        CT thr;
        if ((int)clt <= kSmallCap && U < INFINITY) {
            thr = (CT)U;
        } else {
            thr = block_select<CT, VEC, EPT>(raw_key, n, ksel - 1, (CT)mn, (CT)mx, S, par, ppar);
            if (tid == 0) S.nsmall = 0;
            __syncthreads();
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e)
            if (owned_col<VEC>(e, T, tid) < n && cv[e] < thr) S.small[atomicAdd(&S.nsmall, 1)] = (double)cv[e];
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
    }
    if (topv) {
        const int kout = topk < n ? topk : n;
        float* tv = topv + ((size_t)b * n + row) * (size_t)topk;
        for (int t = tid; t < topk; t += T) tv[t] = t < kout && t < ksel ? (float)S.sorted[t] : INFINITY;
    }

    // ---- median and MAD
    const double med = block_median<CT, VEC, EPT>(raw_key, n, (CT)mn, (CT)mx, S, par, ppar);
    // |c - med| in the storage type: rounding is monotone up to one storage ulp of the largest
    // deviation, far inside the feature tolerance
    const CT med_c = (CT)med;
    auto dev_key = [&](int e) { const CT t = cv[e] - med_c; return t < (CT)0 ? -t : t; };
    CT dhi = (CT)0;
    {
        const CT a1 = (CT)mx - med_c, a2 = med_c - (CT)mn;
        dhi = a1 > a2 ? a1 : a2;
        if (dhi < (CT)0) dhi = (CT)0;
        dhi = dhi + dhi * (CT)1e-6;   // loose upper bound (keys are rounded independently)
    }
    double mad = block_median<CT, VEC, EPT>(dev_key, n, (CT)0, dhi, S, par, ppar);
    if (mad < 1e-9) mad = 1e-9;

    if (tid == 0) {
        float* f = feat + ((size_t)b * n + row) * kFeatDim;
        double gap = 0.0, comp = 0.0, diffi = 0.0;
        if (n >= 2) {
            gap = S.sorted[1] - S.sorted[0];
            comp = gap / ((mx - mn) + 1e-9);
            diffi = 1.0 / ((mx - mn) / (double)(n - 1) + 1e-9);
        }
        const int k10 = n < 10 ? n : 10;
        double km = 0.0;
        for (int q = 0; q < k10; ++q) km += S.sorted[q];
        km /= (double)k10;
        double kv = 0.0;
        for (int q = 0; q < k10; ++q) { const double t = S.sorted[q] - km; kv += t * t; }
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
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double fr = (double)(1 << q);
            const double ang = 2.0 * 3.14159265358979323846 * (double)row * fr / denom;
            f[13 + 2 * q] = (float)sin(ang);
            f[14 + 2 * q] = (float)cos(ang);
        }
    }
}

}  // namespace b200lap
