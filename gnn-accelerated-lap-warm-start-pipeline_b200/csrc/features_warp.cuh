// features_warp.cuh -- row features for short rows (n = 512, 1024, 2048; binary32 storage): ONE WARP PER ROW,
// the row register-resident, no CTA barrier, no cross-warp traffic.
//
// Same algorithm and arithmetic as features_smem.cuh (sample-bracketed exact median / MAD with lane-private
// candidate lists, sign-bit counts, group-minimum bound for the k smallest), reference gnn/features.py:161-243.
// For rows this short the per-row bookkeeping of a multi-warp CTA (cell atomics, leader broadcasts, ~17
// barriers) costs several times the two passes themselves (ncu: 13.6 k warp-instructions per n = 2048 row, 1.7 k
// of them in the passes); with one warp per row every reduction is a `redux`, every "broadcast" a shuffle and
// the passes read registers.  Only the fast path lives here: a row whose bracket misses, whose list or
// candidate buffer overflows, or whose target bin holds more than 64 keys is appended to a redo list that the
// CTA kernel (k_row_features_smem, with its exact fall-backs) processes right after.
#pragma once
#include "features_smem.cuh"

namespace b200lap {

constexpr int kWarpSamp = 512;      // sample keys per row (16 per lane)
constexpr int kWarpKcap = 32;       // lane-private list capacity
constexpr int kWarpCand = 128;
struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

struct __align__(16) FeatWarpScratch {
    int hist[kSelBins];
    float samp[kWarpSamp];
    float list[kWarpKcap * 32];
    float cand[kWarpCand];
    float tiny[kTinyCap2];
    float sorted[kTopKMax];
    int ncand, ntiny;
    float out[4];
    int pad[2];
};

__device__ __forceinline__ float warp_min_ord(float v) { return ord2f(__reduce_min_sync(kFull, f2ord(v))); }
__device__ __forceinline__ float warp_max_ord(float v) { return ord2f(__reduce_max_sync(kFull, f2ord(v))); }
__device__ __forceinline__ int warp_add_i(int v) { return (int)__reduce_add_sync(kFull, (unsigned)v); }

// bracket of two sample ranks by values [L, H] (warp-level twin of bracket_fast); false = give up (redo list)
__device__ __forceinline__ bool bracket_warp(FeatWarpScratch& W, float lo, float hi, int t1, int t2, int heavy, float& L, float& H)
{
    const int lane = lane_id();
    int before = 0;
    for (int level = 0;; ++level) {
        if (!(lo < hi)) { L = lo; H = lo; return true; }
        const float scale = (float)kSelBins / (hi - lo);
        for (int i = lane; i < kWarpSamp; i += 32) {
            const float x = W.samp[i];
            if (x >= lo && x <= hi) atomicAdd(&W.hist[sel_bin256(x, lo, scale)], 1);
        }
        __syncwarp();
        int b1, bb1, c1, b2, bb2, c2;
        find_bin(W.hist, t1 - before, b1, bb1, c1);
        find_bin(W.hist, t2 - before, b2, bb2, c2);
        const int4 z = {0, 0, 0, 0};
        reinterpret_cast<int4*>(W.hist)[2 * lane] = z;
        reinterpret_cast<int4*>(W.hist)[2 * lane + 1] = z;
        __syncwarp();
        const int inside = bb2 + c2 - bb1;
        if (inside <= heavy) {
            const float binw = (hi - lo) * (1.0f / (float)kSelBins);
            L = lo + ((float)b1 - 0.02f) * binw;
            H = lo + ((float)b2 + 1.02f) * binw;
            if (!(L < H)) H = nextafterf(L, INFINITY);
            return true;
        }
        if (level == 2) return false;
        float mn = INFINITY, mx = -INFINITY;
        for (int i = lane; i < kWarpSamp; i += 32) {
            const float x = W.samp[i];
            if (x >= lo && x <= hi) {
                const int b = sel_bin256(x, lo, scale);
                if (b >= b1 && b <= b2) { mn = fminf(mn, x); mx = fmaxf(mx, x); }
            }
        }
        lo = warp_min_ord(mn); hi = warp_max_ord(mx);
        before += bb1;
    }
}

// ranks k1 <= k2 (relative to `below`) of the lane-private lists with keys in [L, Hb]; false = give up
template <bool ABS>
__device__ __forceinline__ bool list_ranks_warp(FeatWarpScratch& W, const ListCursor& cur, float L, float Hb, int q1, int q2, float& a, float& b)
{
    const int lane = lane_id();
    const float scale = (float)kSelBins / (Hb - L);
    walk_list(cur, [&](float x) { atomicAdd(&W.hist[sel_bin256_in(ABS ? fabsf(x) : x, L, scale)], 1); });
    __syncwarp();
    int b1, bb1, c1, b2, bb2, c2;
    find_bin(W.hist, q1, b1, bb1, c1);
    find_bin(W.hist, q2, b2, bb2, c2);
    const int4 z = {0, 0, 0, 0};
    reinterpret_cast<int4*>(W.hist)[2 * lane] = z;
    reinterpret_cast<int4*>(W.hist)[2 * lane + 1] = z;
    const int cnt = b1 == b2 ? c1 : c1 + c2;
    if (cnt > kTinyCap2 || cnt <= 0) { __syncwarp(); return false; }
    if (lane == 0) W.ntiny = 0;
    __syncwarp();
    walk_list(cur, [&](float xr) {
        const float x = ABS ? fabsf(xr) : xr;
        const int bn = sel_bin256_in(x, L, scale);
        if (bn == b1 || bn == b2) { const int q = atomicAdd(&W.ntiny, 1); if (q < kTinyCap2) W.tiny[q] = x; }
    });
    __syncwarp();
    const int k1 = q1 - bb1, k2 = q2 - bb1;
    const float x0 = lane < cnt ? W.tiny[lane] : INFINITY, x1 = lane + 32 < cnt ? W.tiny[lane + 32] : INFINITY;
    int rk0 = 0, rk1 = 0;
    for (int q = 0; q < cnt; ++q) {
        const float o = W.tiny[q];
        rk0 += (o < x0) || (o == x0 && q < lane);
        rk1 += (o < x1) || (o == x1 && q < lane + 32);
    }
    if (lane < cnt && rk0 == k1) W.out[0] = x0;
    if (lane + 32 < cnt && rk1 == k1) W.out[0] = x1;
    if (lane < cnt && rk0 == k2) W.out[1] = x0;
    if (lane + 32 < cnt && rk1 == k2) W.out[1] = x1;
    __syncwarp();
    a = W.out[0]; b = W.out[1];
    __syncwarp();
    return true;
}

template <int EPT>
__global__ void __launch_bounds__(128, 4) k_row_features_warp(FeatSmemArgs a)
{
    B200LAP_DYN_SMEM(smem_raw);
    constexpr int G = EPT / 4;                       // float4 groups per lane
    constexpr int SP = EPT >= 64 ? 1 : (EPT >= 32 ? 2 : 4);   // sample keys per group -> 16 per lane, 512 per row
    const int lane = lane_id(), n = a.n;
    FeatWarpScratch& W = reinterpret_cast<FeatWarpScratch*>(smem_raw)[warp_id()];
    const long long total_rows = (long long)a.batch * n;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (int i = lane; i < kSelBins; i += 32) W.hist[i] = 0;
    __syncwarp();

    const int r1 = (n - 1) >> 1, r2 = n >> 1;
    int ksel = a.topk > 10 ? a.topk : 10;
    if (ksel > kTopKMax) ksel = kTopKMax;
    const int s = kWarpSamp;
    const int sm1 = (int)(((long long)r1 * s) / n), sm2 = (int)(((long long)r2 * s) / n);
    const int st1 = max(0, sm1 - a.delta), st2 = min(s - 1, sm2 + a.delta);
    const int heavy = (st2 - st1) + (st2 - st1) / 2 + s / 32 + 8;
    const double inv_n_d = 1.0 / (double)n;
    const float inv_n = (float)inv_n_d;
    constexpr float kNegLog2e = -1.4426950408889634f;

    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + warp_id(); r < total_rows; r += nwarps) {
        const int b = (int)(r / n), row = (int)(r % n);
        const float4* crow = reinterpret_cast<const float4*>(a.C + (size_t)b * a.inst_stride + (size_t)row * a.ld);
        const float4* cm = reinterpret_cast<const float4*>(a.colmin + (size_t)b * n);
        float c[EPT];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float4 t = __ldcs(crow + g * 32 + lane);
            c[4 * g] = t.x; c[4 * g + 1] = t.y; c[4 * g + 2] = t.z; c[4 * g + 3] = t.w;
        }
        bool redo = false;

        // ---- sample (one, two or four keys of every group, rotated with the lane), its range, the median bracket
        float smn = INFINITY, smx = -INFINITY;
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int u = 0; u < SP; ++u) {
                const int q = SP == 4 ? u : ((lane + g) & (4 / SP - 1)) * SP + u;
                const float x = q == 0 ? c[4 * g] : (q == 1 ? c[4 * g + 1] : (q == 2 ? c[4 * g + 2] : c[4 * g + 3]));
                W.samp[(g * SP + u) * 32 + lane] = x;
                smn = fminf(smn, x); smx = fmaxf(smx, x);
            }
        }
        smn = warp_min_ord(smn); smx = warp_max_ord(smx);
        __syncwarp();
        float L = 0.0f, H = 0.0f;
        if (!bracket_warp(W, smn, smx, st1, st2, heavy, L, H)) redo = true;

        // ---- pass 1
        float tmn = INFINITY, tmx = -INFINITY, tsum = 0.0f;
        int notbest = 0, cless = 0, cgt = 0;
        ListCursor cur;
        cur.init(W.list, lane, 32, 5);
        const bool tie = !(L < H);
        unsigned wlim = (tie || redo) ? 0u : __float_as_uint(H - L) + 1u;
        bool ovf = false;
        auto pass1 = [&](auto tie_tag) {
            constexpr bool TIE = decltype(tie_tag)::value;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 m4 = __ldg(cm + g * 32 + lane);
                const float m[4] = {m4.x, m4.y, m4.z, m4.w};
                if (!TIE && cur.beyond(kWarpKcap - 4)) { wlim = 0u; ovf = true; }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float x = c[4 * g + q];
                    notbest += __float_as_uint(m[q] - x) >> 31;
                    tmn = fminf(tmn, x); tmx = fmaxf(tmx, x);
                    tsum += x;
                    const unsigned db = __float_as_uint(x - L);
                    cless += db >> 31;
                    if (TIE) cgt += __float_as_uint(L - x) >> 31;
                    else if (db < wlim) cur.push(x);
                }
            }
        };
        if (tie) pass1(TrueTag{}); else pass1(FalseTag{});
        const int mycnt = cur.count();
        // upper bound of the ksel-th smallest entry: the ksel-th smallest lane minimum
        float U;
        {
            // rank of every lane minimum by 32 independent shuffle compares (no 16-deep redux chain)
            int rank = 0;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const float o = __shfl_sync(kFull, tmn, q);
                rank += (o < tmn) || (o == tmn && q < lane);
            }
            const unsigned who = __ballot_sync(kFull, rank == ksel - 1);
            U = __shfl_sync(kFull, tmn, __ffs((int)who) - 1);
        }
        const float mn = warp_min_ord(tmn), mx = warp_max_ord(tmx);
        const float sum = warp_sum_f(tsum);
        const int colbest = n - warp_add_i(notbest);
        const int below = warp_add_i(cless);
        const int inside = tie ? n - below - warp_add_i(cgt) : warp_add_i(mycnt);
        if (__any_sync(kFull, ovf)) redo = true;
        const double mean = (double)sum * inv_n_d;
        const float mean_f = (float)mean;

        // ---- median
        float ma = L, mb = L;
        if (!redo) {
            if (!(r1 >= below && r2 < below + inside)) redo = true;
            else if (!tie && !list_ranks_warp<false>(W, cur, L, bracket_upper(L, H - L), r1 - below, r2 - below, ma, mb)) redo = true;
        }
        const double med = ((double)ma + (double)mb) * 0.5;
        const float med_f = (float)med;

        // ---- MAD bracket
        for (int i = lane; i < kWarpSamp; i += 32) W.samp[i] = fabsf(W.samp[i] - med_f);
        float dhi = fmaxf(mx - med_f, med_f - mn);
        if (!(dhi > 0.0f)) dhi = 0.0f;
        __syncwarp();
        float L2 = 0.0f, H2 = 0.0f;
        if (!redo && !bracket_warp(W, 0.0f, dhi, st1, st2, heavy, L2, H2)) redo = true;

        // ---- pass 2
        float tzz = 0.0f, tes = 0.0f, tew = 0.0f;
        int cless2 = 0, cgt2 = 0;
        cur.init(W.list, lane, 32, 5);
        const bool tie2 = !(L2 < H2);
        unsigned wlim2 = (tie2 || redo) ? 0u : __float_as_uint(H2 - L2) + 1u;
        bool ovf2 = false;
        if (lane == 0) W.ncand = 0;
        __syncwarp();
        auto pass2 = [&](auto tie_tag) {
            constexpr bool TIE = decltype(tie_tag)::value;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (!TIE && cur.beyond(kWarpKcap - 4)) { wlim2 = 0u; ovf2 = true; }
                float gmin = INFINITY;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float x = c[4 * g + q];
                    const float z = x - mn;
                    const float ex = exp2_neg_fast(z * kNegLog2e);
                    tzz = fmaf(z, z, tzz);
                    tes += ex;
                    tew = fmaf(ex, z, tew);
                    const float dv = x - med_f;
                    const float key = fabsf(dv);
                    const unsigned db = __float_as_uint(key - L2);
                    cless2 += db >> 31;
                    if (TIE) cgt2 += __float_as_uint(L2 - key) >> 31;
                    else if (db < wlim2) cur.push(dv);
                    gmin = fminf(gmin, x);
                }
                if (gmin < U) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (c[4 * g + q] < U) { const int t = atomicAdd(&W.ncand, 1); if (t < kWarpCand) W.cand[t] = c[4 * g + q]; }
                }
            }
        };
        if (tie2) pass2(TrueTag{}); else pass2(FalseTag{});
        const float zz = warp_sum_f(tzz), es = warp_sum_f(tes), ew = warp_sum_f(tew);
        const int below2 = warp_add_i(cless2);
        const int inside2 = tie2 ? n - below2 - warp_add_i(cgt2) : warp_add_i(cur.count());
        if (__any_sync(kFull, ovf2)) redo = true;
        __syncwarp();
        const int nc = W.ncand;
        if (nc > kWarpCand) redo = true;

        // ---- MAD
        float da = L2, db2 = L2;
        if (!redo) {
            if (!(r1 >= below2 && r2 < below2 + inside2)) redo = true;
            else if (!tie2 && !list_ranks_warp<true>(W, cur, L2, bracket_upper(L2, H2 - L2), r1 - below2, r2 - below2, da, db2)) redo = true;
        }

        if (redo) {
            if (lane == 0) a.redo_list[atomicAdd(a.redo_count, 1)] = (int)r;
            continue;
        }

        // ---- rare extra passes over the registers (uniform): near-best count, exact variance, exp sum without the ones
        const float near_thr = near_threshold(mn, a.torch_mode);
        const double dmean = mean - (double)mn;
        const double var_z = (double)zz * inv_n_d - dmean * dmean;
        const bool need_near = !(near_thr < U);
        const bool need_var = !(var_z * 20.0 > dmean * dmean);
        const bool need_exp = es < 4.0f;
        int nnear = 0, ones = 0;
        float tss = 0.0f, small = 0.0f;
        if (need_near || need_var || need_exp) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const float x = c[e];
                nnear += (x <= near_thr);
                const float dl = x - mean_f;
                tss = fmaf(dl, dl, tss);
                const float z = x - mn;
                if (z > 0.0f) small += exp2_neg_fast(z * kNegLog2e); else ++ones;
            }
            nnear = warp_add_i(nnear); ones = warp_add_i(ones);
            tss = warp_sum_f(tss); small = warp_sum_f(small);
        }

        // ---- the ksel smallest entries (candidates strictly below U, then copies of U), and the finish
        int nearc = 0;
        for (int t = lane; t < (nc > ksel ? nc : ksel); t += 32) {
            if (t < nc) {
                const float mine = W.cand[t];
                nearc += (mine <= near_thr);
                int rank = 0;
                for (int q = 0; q < nc; ++q) {
                    const float o = W.cand[q];
                    rank += (o < mine) || (o == mine && q < t);
                }
                if (rank < ksel) W.sorted[rank] = mine;
            } else {
                W.sorted[t] = U;
            }
        }
        nearc = warp_add_i(nearc);
        __syncwarp();
        float* f = a.feat + ((size_t)b * n + row) * kFeatDim;
        if (lane == 0) {
            const int near = need_near ? nnear : nearc;
            double mad = ((double)da + (double)db2) * 0.5;
            if (mad < 1e-9) mad = 1e-9;
            const float gap = W.sorted[1] - W.sorted[0];
            const float range = mx - mn;                                  // exact difference of two binary32 values, rounded once
            const float comp = __fdividef(gap, range + 1e-9f);
            const float diffi = __fdividef(1.0f, range * (1.0f / (float)(n - 1)) + 1e-9f);
            double km = 0.0;
            for (int q = 0; q < 10; ++q) km += (double)W.sorted[q];
            km *= 0.1;
            double kv = 0.0;
            for (int q = 0; q < 10; ++q) { const double t = (double)W.sorted[q] - km; kv += t * t; }
            float ent;
            if (need_exp) {
                const double es_d = (double)ones + (double)small, sp = es_d + 1e-9;
                ent = (float)((es_d / sp) * log(sp) + (double)ew / sp);
            } else {
                const float sp = es + 1e-9f, inv = __fdividef(1.0f, sp);
                ent = es * inv * logf(sp) + ew * inv;
            }
            const double var = need_var ? (double)tss * inv_n_d : var_z;
            f[0] = mn; f[1] = mx; f[2] = mean_f;
            f[3] = sqrtf((float)(var > 0.0 ? var : 0.0));
            f[4] = (float)mad;
            f[5] = ent;
            f[6] = gap; f[7] = comp;
            f[8] = (float)km;
            f[9] = sqrtf((float)(kv * 0.1));
            f[10] = diffi;
            f[11] = (float)near * inv_n;
            f[12] = (float)colbest * inv_n;
        } else if (lane >= 13 && lane < kFeatDim) {
            f[lane] = a.posenc[(size_t)row * 8 + (lane - 13)];
        }
        if (a.topv && lane < a.topk) a.topv[((size_t)b * n + row) * (size_t)a.topk + lane] = lane < ksel ? W.sorted[lane] : INFINITY;
        __syncwarp();
    }
}

}  // namespace b200lap
