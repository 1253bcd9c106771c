// features.cuh -- the 21-D row features and the per-row top-k selection in ONE read of C.
//
// Reference: gnn/features.py:161-243 (compute_row_features), :21-31 (_positional_encodings);
// the top-k values feed gnn/one_gnn.py:143-147 (topk of cost - u_pre: only the VALUES are used,
// and subtracting a per-row constant is monotone, so the k smallest raw entries are selected
// here, before the MLP runs -- SURVEY.md App. C.2).
//
// One CTA per row; every thread keeps its EPT entries of the row in registers (VEC-interleaved,
// 128-bit loads), so HBM is read once and every later "pass" is register-resident.  Exact order
// statistics (median, MAD, k-th smallest) come from a value-linear histogram select: bins are
// linear between the current lower/upper bound (monotone in the value, so bins are ordered),
// the bin holding the wanted rank is either small enough to rank exhaustively or becomes the
// next, narrower range.  Statistics are accumulated in binary64; exp/log use the SFU in binary32
// (tolerance 1e-4 relative, stated in the tests).
#pragma once
#include "common.cuh"
#include "frontend.cuh"   // RowLoad, owned_col

namespace b200lap {

constexpr int kFeatDim = 21;
constexpr int kSelBins = 1024;
constexpr int kSelCap = 256;
constexpr int kTopKMax = 32;

struct FeatShared {
    double red[2][4][32];
    int hist[kSelBins];
    int wsum[32];
    double cand[kSelCap];
    double sorted[kTopKMax];
    int ncand;
    int bin, before, count;
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
    return t >= (KT)(kSelBins - 1) ? kSelBins - 1 : (int)t;
}

// Exact r-th smallest (0-based) of the valid keys, all of which lie in [lo, hi].
template <typename KT, int VEC, int EPT, typename KeyF>
__device__ KT block_select(KeyF kf, int n, int r, KT lo, KT hi, FeatShared& S, int& par)
{
    const int T = blockDim.x, tid = threadIdx.x;
    KT key[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) key[e] = kf(e);
    while (true) {
        if (!(lo < hi)) return lo;
        for (int b = tid; b < kSelBins; b += T) S.hist[b] = 0;
        if (tid == 0) S.ncand = 0;
        __syncthreads();
        const KT scale = (KT)kSelBins / (hi - lo);
#pragma unroll
        for (int e = 0; e < EPT; ++e)
            if (owned_col<VEC>(e, T, tid) < n && key[e] >= lo && key[e] <= hi) atomicAdd(&S.hist[sel_bin(key[e], lo, scale)], 1);
        __syncthreads();
        // locate the bin holding rank r: blocked ownership of bins, two-level exclusive scan
        const int bpt = (kSelBins + T - 1) / T;
        const int b0 = tid * bpt;
        int local = 0;
        for (int q = 0; q < bpt; ++q)
            if (b0 + q < kSelBins) local += S.hist[b0 + q];
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane_id() >= o) incl += t;
        }
        if (lane_id() == 31) S.wsum[warp_id()] = incl;
        __syncthreads();
        int wbase = 0;
        {
            const int t = lane_id() < warp_id() ? S.wsum[lane_id()] : 0;
            wbase = warp_sum_i(t);
        }
        int run = wbase + incl - local;
        if (r >= run && r < run + local) {
            for (int q = 0; q < bpt && b0 + q < kSelBins; ++q) {
                const int h = S.hist[b0 + q];
                if (r < run + h) { S.bin = b0 + q; S.before = run; S.count = h; break; }
                run += h;
            }
        }
        __syncthreads();
        const int bin = S.bin, before = S.before, count = S.count;
        if (count <= kSelCap) {
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if (owned_col<VEC>(e, T, tid) < n && key[e] >= lo && key[e] <= hi && sel_bin(key[e], lo, scale) == bin)
                    S.cand[atomicAdd(&S.ncand, 1)] = (double)key[e];
            __syncthreads();
            const int want = r - before;
            for (int t = tid; t < count; t += T) {
                const double mine = S.cand[t];
                int rank = 0;
                for (int q = 0; q < count; ++q) {
                    const double o = S.cand[q];
                    rank += (o < mine) || (o == mine && q < t);
                }
                if (rank == want) S.result = mine;
            }
            __syncthreads();
            return (KT)S.result;
        }
        // too many keys in the bin: shrink the range to the bin's own extent and go again
        double bmin = INFINITY, bmax = -INFINITY, z0 = 0.0, z1 = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; ++e)
            if (owned_col<VEC>(e, T, tid) < n && key[e] >= lo && key[e] <= hi && sel_bin(key[e], lo, scale) == bin) {
                bmin = (double)key[e] < bmin ? (double)key[e] : bmin;
                bmax = (double)key[e] > bmax ? (double)key[e] : bmax;
            }
        block_red4<OP_MIN, OP_MAX, OP_SUM, OP_SUM>(S, par, bmin, bmax, z0, z1);
        r -= before;
        lo = (KT)bmin;
        hi = (KT)bmax;
    }
}

// median = mean of the order statistics (n-1)/2 and n/2 (numpy's definition)
template <typename KT, int VEC, int EPT, typename KeyF>
__device__ double block_median(KeyF kf, int n, KT lo, KT hi, FeatShared& S, int& par)
{
    const int T = blockDim.x, tid = threadIdx.x;
    const int r1 = (n - 1) / 2;
    const KT a = block_select<KT, VEC, EPT>(kf, n, r1, lo, hi, S, par);
    if (n & 1) return (double)a;
    double le = 0.0, above = INFINITY, z0 = 0.0, z1 = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e)
        if (owned_col<VEC>(e, T, tid) < n) {
            const KT ke = kf(e);
            if (ke <= a) le += 1.0;
            else above = (double)ke < above ? (double)ke : above;
        }
    block_red4<OP_SUM, OP_MIN, OP_SUM, OP_SUM>(S, par, le, above, z0, z1);
    const double b = ((int)le > r1 + 1) ? (double)a : above;
    return ((double)a + b) / 2.0;
}

template <typename CT, int VEC, int EPT>
__global__ void __launch_bounds__(1024) k_row_features(
    const CT* __restrict__ C, long long inst_stride, int ld, int n, int topk,
    const CT* __restrict__ colmin /* [B][n] */, float* __restrict__ feat /* [B][n][21] */,
    float* __restrict__ topv /* [B][n][topk] or null */)
{
    __shared__ FeatShared S;
    const int b = blockIdx.y, row = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
    const CT* crow = C + (size_t)b * inst_stride + (size_t)row * ld;
    const CT* cm = colmin + (size_t)b * n;
    int par = 0;
    CT cv[EPT];
    int colbest = 0;
    double mn = INFINITY, mx = -INFINITY, sum = 0.0, z = 0.0;
#pragma unroll
    for (int g = 0; g < EPT / VEC; ++g) {
        const int col = owned_col<VEC>(g * VEC, T, tid);
        if (col < n) {
            RowLoad<CT, VEC>::ld(crow + col, &cv[g * VEC]);
            CT mm[VEC];
            RowLoad<CT, VEC>::ld(cm + col, mm);
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                const double c = (double)cv[g * VEC + q];
                colbest += (cv[g * VEC + q] == mm[q]);
                mn = c < mn ? c : mn;
                mx = c > mx ? c : mx;
                sum += c;
            }
        } else {
#pragma unroll
            for (int q = 0; q < VEC; ++q) cv[g * VEC + q] = (CT)0;
        }
    }
    block_red4<OP_MIN, OP_MAX, OP_SUM, OP_SUM>(S, par, mn, mx, sum, z);
    const double mean = sum / (double)n;
    const double near_thr = mn * 1.1;
    double ssq = 0.0, esum = 0.0, near = 0.0, cb = (double)colbest;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        if (owned_col<VEC>(e, T, tid) < n) {
            const double c = (double)cv[e];
            const double dlt = c - mean;
            ssq += dlt * dlt;
            esum += (double)__expf(-(float)(c - mn));
            near += (c <= near_thr) ? 1.0 : 0.0;
        }
    }
    block_red4<OP_SUM, OP_SUM, OP_SUM, OP_SUM>(S, par, ssq, esum, near, cb);
    const float inv = (float)(1.0 / (esum + 1e-9));
    double ent = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e)
        if (owned_col<VEC>(e, T, tid) < n) {
            const float p = __expf(-(float)((double)cv[e] - mn)) * inv;
            ent -= (double)(p * __logf(p + 1e-9f));
        }
    block_red4<OP_SUM, OP_SUM, OP_SUM, OP_SUM>(S, par, ent, z1, z2, z3);

    // ---- k smallest, ascending (k = what the features (10) and the model (topk) need)
    int ksel = topk > 10 ? topk : 10;
    if (ksel > n) ksel = n;
    if (ksel > kTopKMax) ksel = kTopKMax;
    auto raw_key = [&](int e) { return cv[e]; };
    const CT thr = block_select<CT, VEC, EPT>(raw_key, n, ksel - 1, (CT)mn, (CT)mx, S, par);
    if (tid == 0) S.ncand = 0;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPT; ++e)
        if (owned_col<VEC>(e, T, tid) < n && cv[e] < thr) S.cand[atomicAdd(&S.ncand, 1)] = (double)cv[e];
    __syncthreads();
    {
        const int c = S.ncand;   // < ksel
        for (int t = tid; t < ksel; t += T) {
            if (t < c) {
                const double mine = S.cand[t];
                int rank = 0;
                for (int q = 0; q < c; ++q) {
                    const double o = S.cand[q];
                    rank += (o < mine) || (o == mine && q < t);
                }
                S.sorted[rank] = mine;
            } else {
                S.sorted[t] = (double)thr;
            }
        }
    }
    __syncthreads();
    if (topv) {
        const int kout = topk < n ? topk : n;
        float* tv = topv + ((size_t)b * n + row) * (size_t)topk;
        for (int t = tid; t < topk; t += T) tv[t] = t < kout && t < ksel ? (float)S.sorted[t] : INFINITY;
    }

    // ---- median and MAD
    const double med = block_median<CT, VEC, EPT>(raw_key, n, (CT)mn, (CT)mx, S, par);
    // |c - med| rounded to the storage type: rounding is monotone, so it commutes with the order
    // statistics and costs at most one storage-type ulp on the result
    auto dev_key = [&](int e) { return (CT)fabs((double)cv[e] - med); };
    double dmn = INFINITY, dmx = -INFINITY;
    z1 = 0.0; z2 = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        if (owned_col<VEC>(e, T, tid) < n) {
            const double dk = (double)dev_key(e);
            dmn = dk < dmn ? dk : dmn;
            dmx = dk > dmx ? dk : dmx;
        }
    }
    block_red4<OP_MIN, OP_MAX, OP_SUM, OP_SUM>(S, par, dmn, dmx, z1, z2);
    double mad = block_median<CT, VEC, EPT>(dev_key, n, (CT)dmn, (CT)dmx, S, par);
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
