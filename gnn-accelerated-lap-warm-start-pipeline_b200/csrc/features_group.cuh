// features_group.cuh -- the fast path of the 21-D row features + per-row top-k for binary32 storage and
// n = 32 * G * EPL (512 ... 16384):  a GROUP of G warps owns one row at a time, a persistent CTA holds
// 16 / G groups, every group has its own row buffer in shared memory.
//
// Reference: gnn/features.py:161-243 (compute_row_features); gnn/one_gnn.py:143-147 (top-k values).
// Same arithmetic as features_warp.cuh / features_smem.cuh (those remain the exact fall-backs: a row this
// kernel gives up on -- bracket miss, list overflow, tie-heavy target bin, too many top-k candidates -- is
// appended to a redo list that k_row_features_smem processes right after).
//
// What is different, and why (ncu of the round-1 kernels: both were bound by the SHARED-ATOMIC pipe -- a
// full-warp ATOMS costs ~64 cycles of the SM's load/store unit and the sample / list histograms issued
// 50 (n = 2048) to 300 (n = 16384) of them per row):
//   * no histogram atomics at all.  Every warp SORTS its own 256-key strided sample of the row in registers
//     (bitonic network, 8 keys per lane: 408 instructions, 120 shuffles); the median bracket [L, H] is a
//     pair of sample order statistics (mean over the group's warps), the top-k bound comes from the sorted
//     lane minima, the MAD bracket from ONE bitonic merge of |sample - median| (a V-shaped, i.e. bitonic,
//     sequence).  Distribution free: no value-linear bins, no heavy-bracket re-levels.
//   * the exact rank inside a bracket list is found with LANE-PRIVATE BYTE HISTOGRAMS (64 bins x 32 lanes,
//     plain LDS/STS read-modify-write, no conflicts), summed with dp4a, then a <= 32-key warp sort.
//   * the row lives in shared memory, not registers: ~16 resident warps per SM and the NEXT row's segment
//     is requested (cp.async.bulk, one per warp) the moment the warp has finished its last pass, so the
//     copy overlaps the list selection, the finish and the other groups.
//   * all 16 warps of an SM stay busy at every n: G = 1 (n <= 2048), 2 (4096), 4 (8192), 8 (16384);
//     groups synchronise with named barriers (bar.sync id, 32 G), never with __syncthreads.
#pragma once
#include "features_warp.cuh"

namespace b200lap {

constexpr int kGrpSamp = 256;     // sample keys per warp (8 per lane)
constexpr int kGrpBinsMax = 128;  // bins of the lane-private byte histograms: 64, or 128 for rows of >= 128 entries per lane
constexpr int kGrpTiny = 64;
constexpr int kGrpCand = 128;
constexpr int kGrpPart = 12;

template <int G, int BINS>
struct __align__(16) GrpShared {
    float part[G][kGrpPart];
    int cnts[G][BINS];
    float tiny[kGrpTiny];
    float cand[kGrpCand];
    float sorted[32];
    int ntiny, ncand;
    float res[2];
};

__host__ __device__ constexpr int grp_parities(int G) { return G > 1 ? 2 : 1; }
__host__ __device__ constexpr int grp_bins(int EPL) { return EPL >= 128 ? 128 : 64; }
// bytes of dynamic shared memory for Q groups of G warps, rows of n floats (none when the row is streamed from
// global memory instead of being staged), `bins` histogram bins, list capacity kcap per lane
template <int G>
__host__ __device__ constexpr size_t grp_smem_bytes(int Q, int n, int kcap, int bins, bool stream) {
    return 128 + (size_t)Q * grp_parities(G) * (bins == 128 ? sizeof(GrpShared<G, 128>) : sizeof(GrpShared<G, 64>)) +
           (size_t)Q * G * ((size_t)bins * 32 + (size_t)kcap * 128) + (stream ? 0 : (size_t)Q * n * 4);
}

// ---- group barrier (G warps) ------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void grp_sync(int q) {
    __syncwarp();                 // bar.sync is warp-aligned: the lanes must have reconverged (list walks have per-lane trip counts)
    if (G == 1) return;
#ifdef B200LAP_EMUL
    emul::named_barrier(1 + q, 32 * G);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * G) : "memory");
#endif
}

// ---- bitonic networks over 32 R keys, key e = lane * R + r, ascending -----------------------------------
template <int R>
__device__ __forceinline__ void grp_cross(float (&x)[R], int lane_mask, bool lower, bool mirror) {
    float p[R];
#pragma unroll
    for (int r = 0; r < R; ++r) p[r] = __shfl_xor_sync(kFull, mirror ? x[R - 1 - r] : x[r], lane_mask);
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = lower ? fminf(x[r], p[r]) : fmaxf(x[r], p[r]);
}
template <int R>
__device__ __forceinline__ void grp_local(float (&x)[R], int m) {   // compare-exchange r with r ^ m (m a mask of low bits)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int p = r ^ m;
        if (p > r) { const float lo = fminf(x[r], x[p]), hi = fmaxf(x[r], x[p]); x[r] = lo; x[p] = hi; }
    }
}
template <int R> struct Log2R { static constexpr int v = R == 1 ? 0 : (R == 2 ? 1 : (R == 4 ? 2 : 3)); };

// the merge steps j = from .. 0 of a block (compare e with e ^ 2^j, the lower index keeps the minimum)
template <int R>
__device__ __forceinline__ void grp_merge_steps(float (&x)[R], int from) {
    constexpr int LR = Log2R<R>::v;
    const int lane = lane_id();
#pragma unroll
    for (int j = 9; j >= 0; --j) {
        if (j > from) continue;
        if (j < LR) grp_local<R>(x, 1 << j);
        else grp_cross<R>(x, 1 << (j - LR), (lane & (1 << (j - LR))) == 0, false);
    }
}
template <int R>
__device__ __forceinline__ void warp_bitonic_sort(float (&x)[R]) {
    constexpr int LR = Log2R<R>::v;
    const int lane = lane_id();
#pragma unroll
    for (int k = 1; k <= LR + 5; ++k) {
        // mirror step: e with e ^ (2^k - 1)
        if (k <= LR) grp_local<R>(x, (1 << k) - 1);
        else grp_cross<R>(x, (1 << (k - LR)) - 1, ((lane >> (k - LR - 1)) & 1) == 0, true);
        if (k >= 2) grp_merge_steps<R>(x, k - 2);
    }
}
// sorts a bitonic sequence (here: V-shaped) ascending
template <int R>
__device__ __forceinline__ void warp_bitonic_merge(float (&x)[R]) { grp_merge_steps<R>(x, Log2R<R>::v + 4); }

// key of rank t (0 .. 32 R - 1) of a sorted register array
template <int R>
__device__ __forceinline__ float sorted_at(const float (&x)[R], int t) {
    constexpr int LR = Log2R<R>::v;
    float v = x[0];
#pragma unroll
    for (int r = 1; r < R; ++r) v = (t & (R - 1)) == r ? x[r] : v;
    return __shfl_sync(kFull, v, t >> LR);
}

__device__ __forceinline__ int dp4a_ones(unsigned w, int acc) {
#ifdef B200LAP_EMUL
    return acc + (int)(w & 255u) + (int)((w >> 8) & 255u) + (int)((w >> 16) & 255u) + (int)(w >> 24);
#else
    return (int)__dp4a(w, 0x01010101u, (unsigned)acc);
#endif
}

// Exact keys of the list ranks k1 <= k2 = k1 + {0, 1} of the group's lane-private lists (keys in [lo, hi]).
// Every warp of the group calls it and gets the same answer; false = give up (group-uniform).
template <int G, int BINS>
__device__ __forceinline__ bool grp_select(GrpShared<G, BINS>& S, unsigned char* bh, const float* list, const ListCursor& cur, int q, int w, float lo, float hi,
                                           int k1, int k2, float& ra, float& rb)
{
    constexpr int NB = BINS / 32;                     // bins per lane: lane owns bins lane, lane + 32, ...
    const int lane = lane_id();
    const float scale = (float)BINS / (hi - lo);
    {
        const uint4 z = {0u, 0u, 0u, 0u};
        uint4* b4 = reinterpret_cast<uint4*>(bh);
#pragma unroll
        for (int i = 0; i < BINS / 16; ++i) b4[i * 32 + lane] = z;
    }
    if (w == 0 && lane == 0) S.ntiny = 0;
    __syncwarp();
    auto bin_of = [&](float x) { const int b = (int)((x - lo) * scale); return b > BINS - 1 ? BINS - 1 : b; };
    const int mycnt = cur.count(), maxcnt = warp_max_i(mycnt);
    const float* mylist = list + lane;
    for (int t = 0; t < maxcnt; ++t) {
        if (t < mycnt) {
            unsigned char* h = bh + bin_of(mylist[t * 32]) * 32 + lane;
            *h = (unsigned char)(*h + 1);
        }
    }
    __syncwarp();
    int c[NB], inc[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const uint4* r0 = reinterpret_cast<const uint4*>(bh + (lane + 32 * k) * 32);
        const uint4 a0 = r0[0], a1 = r0[1];
        int t = 0;
        t = dp4a_ones(a0.x, t); t = dp4a_ones(a0.y, t); t = dp4a_ones(a0.z, t); t = dp4a_ones(a0.w, t);
        t = dp4a_ones(a1.x, t); t = dp4a_ones(a1.y, t); t = dp4a_ones(a1.z, t); t = dp4a_ones(a1.w, t);
        c[k] = t;
    }
    if (G > 1) {
#pragma unroll
        for (int k = 0; k < NB; ++k) S.cnts[w][lane + 32 * k] = c[k];
        grp_sync<G>(q);
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            int t = 0;
#pragma unroll
            for (int g = 0; g < G; ++g) t += S.cnts[g][lane + 32 * k];
            c[k] = t;
        }
    }
    // inclusive scans in bin order (segment k holds bins 32 k .. 32 k + 31)
    int carry = 0;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        int v = c[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, v, o);
            if (lane >= o) v += t;
        }
        inc[k] = v + carry;
        carry = __shfl_sync(kFull, inc[k], 31);
    }
    auto locate = [&](int k, int& b, int& before, int& cnt) {
        b = BINS - 1; before = 0; cnt = 0;
        bool found = false;
#pragma unroll
        for (int sgm = 0; sgm < NB; ++sgm) {
            const unsigned m = __ballot_sync(kFull, k < inc[sgm]);
            if (!found && m != 0u) {
                const int src = __ffs((int)m) - 1;
                cnt = __shfl_sync(kFull, c[sgm], src);
                before = __shfl_sync(kFull, inc[sgm], src) - cnt;
                b = 32 * sgm + src;
                found = true;
            }
        }
    };
    int b1, bb1, n1, b2, bb2, n2;
    locate(k1, b1, bb1, n1);
    locate(k2, b2, bb2, n2);
    const int cnt = b1 == b2 ? n1 : n1 + n2;
    if (__any_sync(kFull, cnt > kGrpTiny || cnt <= 0)) return false;     // (the vote tells the compiler the branch is warp-uniform)
    for (int t = 0; t < maxcnt; ++t) {
        if (t < mycnt) {
            const float x = mylist[t * 32];
            const int b = bin_of(x);
            if (b == b1 || b == b2) { const int p = atomicAdd(&S.ntiny, 1); if (p < kGrpTiny) S.tiny[p] = x; }
        }
    }
    grp_sync<G>(q);
    if (G == 1 || w == 0) {
        float v[2] = {2 * lane < cnt ? S.tiny[2 * lane] : INFINITY, 2 * lane + 1 < cnt ? S.tiny[2 * lane + 1] : INFINITY};
        warp_bitonic_sort<2>(v);
        const int x1 = k1 - bb1, x2 = b1 == b2 ? k2 - bb1 : n1 + (k2 - bb2);
        ra = sorted_at<2>(v, x1);
        rb = sorted_at<2>(v, x2);
        if (G > 1 && lane == 0) { S.res[0] = ra; S.res[1] = rb; }
    }
    if (G > 1) {
        grp_sync<G>(q);
        ra = S.res[0]; rb = S.res[1];
    }
    return true;
}

// STREAM: the row is not staged in shared memory; pass 1 reads it from global memory (DRAM), pass 2 and the rare extra
// pass read it again (L2: the row is 32 G EPL x 4 bytes and was read a moment ago).  That frees the 64 KB a row of
// n = 16384 takes, so such rows can be owned by FEWER warps (the per-warp bookkeeping -- sample sort, list selections,
// reductions -- is paid per warp, not per entry) while 16 warps stay resident.
template <int G, int EPL, bool STREAM>
__global__ void __launch_bounds__(512, 1) k_row_features_group(FeatSmemArgs a)
{
    B200LAP_DYN_SMEM(smem_raw);
    constexpr int BINS = grp_bins(EPL);
    constexpr int NG = EPL / 4;                          // float4 groups per lane
    constexpr int P = grp_parities(G);
    constexpr int SEG = EPL * 32;                        // floats per warp segment
    const int lane = lane_id(), n = a.n;
    const int Q = (int)(blockDim.x >> 5) / G;            // groups in this CTA
    const int q = warp_id() / G, w = warp_id() % G;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    GrpShared<G, BINS>* SH = reinterpret_cast<GrpShared<G, BINS>*>(smem_raw + 128) + (size_t)q * P;
    unsigned char* warp_base = smem_raw + 128 + (size_t)Q * P * sizeof(GrpShared<G, BINS>) + (size_t)warp_id() * ((size_t)BINS * 32 + (size_t)a.kcap * 128);
    unsigned char* bh = warp_base;
    float* list = reinterpret_cast<float*>(warp_base + BINS * 32);
    float* rowbuf = reinterpret_cast<float*>(smem_raw + 128 + (size_t)Q * P * sizeof(GrpShared<G, BINS>) + (size_t)Q * G * ((size_t)BINS * 32 + (size_t)a.kcap * 128)) + (STREAM ? 0 : (size_t)q * n);
    const float4* seg4 = reinterpret_cast<const float4*>(rowbuf + (size_t)w * SEG);

    const long long total_rows = (long long)a.batch * n;
    const long long ngroups = (long long)gridDim.x * Q;
    const int r1 = (n - 1) >> 1, r2 = n >> 1;
    int ksel = a.topk > 10 ? a.topk : 10;
    if (ksel > kTopKMax) ksel = kTopKMax;
    const int t1 = kGrpSamp / 2 - 1 - a.delta, t2 = kGrpSamp / 2 + a.delta;     // local sample ranks of the bracket
    const double inv_n_d = 1.0 / (double)n;
    const float inv_n = (float)inv_n_d;
    constexpr float kNegLog2e = -1.4426950408889634f;
    constexpr float invG = 1.0f / (float)G;
    constexpr int STRIDE = EPL / 8;                      // row entries per pooled sample key

    auto row_src = [&](long long r) { return a.C + (r / n) * a.inst_stride + (r % n) * (long long)a.ld + (size_t)w * SEG; };
#ifndef B200LAP_EMUL
    if (!STREAM && threadIdx.x == 0) {
        for (int i = 0; i < Q; ++i) mbar_init(&full[i], G);
        mbar_fence_init();
    }
    __syncthreads();
    // request this warp's segment of row r (the buffer segment must no longer be read by anybody)
    auto request = [&](long long r) {
        if (STREAM) return;
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(&full[q], (unsigned)SEG * 4u);
            bulk_g2s(rowbuf + (size_t)w * SEG, row_src(r), (unsigned)SEG * 4u, &full[q]);
        }
    };
    unsigned phase = 0u;
#else
    auto request = [&](long long r) {
        if (STREAM) return;
        __syncwarp();
        const float* src = row_src(r);
        float* dst = rowbuf + (size_t)w * SEG;
        for (int i = lane; i < SEG; i += 32) dst[i] = src[i];
    };
#endif

    long long r = (long long)blockIdx.x * Q + q;
    if (r < total_rows) request(r);
    int par = 0;
    for (; r < total_rows; r += ngroups) {
        const long long rnext = r + ngroups;
        GrpShared<G, BINS>& S = SH[P == 2 ? par : 0];
        par ^= 1;
#ifndef B200LAP_EMUL
        if (!STREAM) { mbar_wait(&full[q], phase); phase ^= 1u; }
#else
        if (!STREAM) grp_sync<G>(q);
#endif
        const int b = (int)(r / n), row = (int)(r % n);
        const float* grow = a.C + (size_t)b * a.inst_stride + (size_t)row * a.ld;             // the row in global memory
        const float4* gseg4 = reinterpret_cast<const float4*>(grow + (size_t)w * SEG);
        auto ld_pass1 = [&](int idx) { return STREAM ? __ldg(gseg4 + idx) : seg4[idx]; };     // DRAM: keep the line for pass 2
        auto ld_again = [&](int idx) { return STREAM ? __ldcs(gseg4 + idx) : seg4[idx]; };    // L2 hit, last use
        const float4* cm = reinterpret_cast<const float4*>(a.colmin + (size_t)b * n + (size_t)w * SEG);
        bool redo = false;

        // ---- this warp's sample: pooled key i = t * G + w (t = j * 32 + lane), entry i * STRIDE + a lane-dependent skew
        float sx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = (j * 32 + lane) * G + w;
            const int pos = i * STRIDE + ((lane >> 2) & (STRIDE - 1));
            sx[j] = STREAM ? __ldg(grow + pos) : rowbuf[pos];
        }
        warp_bitonic_sort<8>(sx);
        float L = sorted_at<8>(sx, t1), H = sorted_at<8>(sx, t2);
        if (G > 1) {
            if (lane == 0) { S.part[w][0] = L; S.part[w][1] = H; }
            grp_sync<G>(q);
            L = 0.0f; H = 0.0f;
#pragma unroll
            for (int g = 0; g < G; ++g) { L += S.part[g][0]; H += S.part[g][1]; }
            L *= invG; H *= invG;
            grp_sync<G>(q);          // part[] is rewritten after pass 1
        }

        // ---- pass 1: min, max, sum, is_col_best, lane minimum; median bracket (count below L, list of [L, H])
        float tmn = INFINITY, tmx = -INFINITY, tsum = 0.0f;
        int notbest = 0, cless = 0, cgt = 0;
        ListCursor cur;
        cur.init(list, lane, 32, 5);
        const bool tie = !(L < H);
        unsigned wlim = tie ? 0u : __float_as_uint(H - L) + 1u;
        bool ovf = false;
        auto pass1 = [&](auto tie_tag) {
            constexpr bool TIE = decltype(tie_tag)::value;
#pragma unroll 4
            for (int g = 0; g < NG; ++g) {
                const float4 c4 = ld_pass1(g * 32 + lane);
                const float4 m4 = __ldg(cm + g * 32 + lane);
                const float c[4] = {c4.x, c4.y, c4.z, c4.w};
                const float m[4] = {m4.x, m4.y, m4.z, m4.w};
                if (!TIE && cur.beyond(a.kcap - 4)) { wlim = 0u; ovf = true; }
                tmn = fminf(fminf(tmn, fminf(c[0], c[1])), fminf(c[2], c[3]));      // FMNMX3 pairs
                tmx = fmaxf(fmaxf(tmx, fmaxf(c[0], c[1])), fmaxf(c[2], c[3]));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x = c[e];
                    notbest += __float_as_uint(m[e] - x) >> 31;
                    tsum += x;
                    const unsigned db = __float_as_uint(x - L);
                    cless += db >> 31;
                    if (TIE) cgt += __float_as_uint(L - x) >> 31;
                    else if (db < wlim) cur.push(x);
                }
            }
        };
        if (__all_sync(kFull, tie)) pass1(TrueTag{}); else pass1(FalseTag{});
        // upper bound of the ksel-th smallest entry: every warp has at least ceil(ksel / G) entries at or below its
        // ceil(ksel / G)-th smallest lane minimum, so the maximum of those over the group's warps bounds ksel entries
        float U;
        {
            float lm[1] = {tmn};
            warp_bitonic_sort<1>(lm);
            U = __shfl_sync(kFull, lm[0], (ksel + G - 1) / G - 1);
        }
        float mn = warp_min_ord(tmn), mx = warp_max_ord(tmx);
        float sum = warp_sum_f(tsum);
        int nb = warp_add_i(notbest), below = warp_add_i(cless);
        int inside = tie ? SEG - below - warp_add_i(cgt) : warp_add_i(cur.count());
        bool anyovf = __any_sync(kFull, ovf);
        if (G > 1) {
            if (lane == 0) {
                float* pp = S.part[w];
                pp[0] = mn; pp[1] = mx; pp[2] = sum; pp[3] = U;
                pp[4] = __int_as_float(nb); pp[5] = __int_as_float(below); pp[6] = __int_as_float(inside); pp[7] = __int_as_float(anyovf ? 1 : 0);
            }
            grp_sync<G>(q);
            mn = INFINITY; mx = -INFINITY; sum = 0.0f; U = -INFINITY; nb = 0; below = 0; inside = 0; anyovf = false;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float* pp = S.part[g];
                mn = fminf(mn, pp[0]); mx = fmaxf(mx, pp[1]); sum += pp[2]; U = fmaxf(U, pp[3]);
                nb += __float_as_int(pp[4]); below += __float_as_int(pp[5]); inside += __float_as_int(pp[6]); anyovf |= __float_as_int(pp[7]) != 0;
            }
        }
        const int colbest = n - nb;
        int why = 0;
        if (anyovf) { redo = true; why = 1; }
        const double mean = (double)sum * inv_n_d;
        const float mean_f = (float)mean;

        // ---- median
        float ma = L, mb = L;
        if (!redo && !(r1 >= below && r2 < below + inside)) { redo = true; why = 2; }
        if (!__any_sync(kFull, redo) && !__all_sync(kFull, tie)) {
            if (!grp_select<G, BINS>(S, bh, list, cur, q, w, L, bracket_upper(L, H - L), r1 - below, r2 - below, ma, mb)) { redo = true; why = 3; }
        }
        if (__any_sync(kFull, redo)) {
            // the row buffer is free: get the next row, hand this one to the CTA kernel
            if (rnext < total_rows) request(rnext);
            if (w == 0 && lane == 0) { a.redo_list[atomicAdd(a.redo_count, 1)] = (int)r; atomicAdd(a.redo_count + why, 1); }
            continue;
        }
        const double med = ((double)ma + (double)mb) * 0.5;
        const float med_f = (float)med;

        // ---- MAD bracket: |sample - median| is V-shaped along the sorted sample, one bitonic merge sorts it
#pragma unroll
        for (int j = 0; j < 8; ++j) sx[j] = fabsf(sx[j] - med_f);
        warp_bitonic_merge<8>(sx);
        float L2 = sorted_at<8>(sx, t1), H2 = sorted_at<8>(sx, t2);
        if (G > 1) {
            grp_sync<G>(q);          // everybody has read the pass-1 partials
            if (lane == 0) { S.part[w][0] = L2; S.part[w][1] = H2; }
            if (w == 0 && lane == 0) S.ncand = 0;
            grp_sync<G>(q);
            L2 = 0.0f; H2 = 0.0f;
#pragma unroll
            for (int g = 0; g < G; ++g) { L2 += S.part[g][0]; H2 += S.part[g][1]; }
            L2 *= invG; H2 *= invG;
            grp_sync<G>(q);
        } else {
            if (lane == 0) S.ncand = 0;
            __syncwarp();
        }

        // ---- pass 2: variance and exp sums around the exact minimum, top-k candidates below U, MAD bracket
        float tzz = 0.0f, tes = 0.0f, tew = 0.0f;
        int cless2 = 0, cgt2 = 0;
        cur.init(list, lane, 32, 5);
        const bool tie2 = !(L2 < H2);
        unsigned wlim2 = tie2 ? 0u : __float_as_uint(H2 - L2) + 1u;
        bool ovf2 = false;
        auto pass2 = [&](auto tie_tag) {
            constexpr bool TIE = decltype(tie_tag)::value;
#pragma unroll 4
            for (int g = 0; g < NG; ++g) {
                const float4 c4 = ld_again(g * 32 + lane);
                const float c[4] = {c4.x, c4.y, c4.z, c4.w};
                if (!TIE && cur.beyond(a.kcap - 4)) { wlim2 = 0u; ovf2 = true; }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x = c[e];
                    const float z = x - mn;
                    const float ex = exp2_neg_fast(z * kNegLog2e);
                    tzz = fmaf(z, z, tzz);
                    tes += ex;
                    tew = fmaf(ex, z, tew);
                    const float key = fabsf(x - med_f);
                    const unsigned db = __float_as_uint(key - L2);
                    cless2 += db >> 31;
                    if (TIE) cgt2 += __float_as_uint(L2 - key) >> 31;
                    else if (db < wlim2) cur.push(key);
                }
                if (fminf(fminf(c[0], c[1]), fminf(c[2], c[3])) < U) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (c[e] < U) { const int t = atomicAdd(&S.ncand, 1); if (t < kGrpCand) S.cand[t] = c[e]; }
                }
            }
        };
        if (__all_sync(kFull, tie2)) pass2(TrueTag{}); else pass2(FalseTag{});
        float zz = warp_sum_f(tzz), es = warp_sum_f(tes), ew = warp_sum_f(tew);
        int below2 = warp_add_i(cless2);
        int inside2 = tie2 ? SEG - below2 - warp_add_i(cgt2) : warp_add_i(cur.count());
        bool anyovf2 = __any_sync(kFull, ovf2);
        if (G > 1) {
            if (lane == 0) {
                float* pp = S.part[w];
                pp[0] = zz; pp[1] = es; pp[2] = ew;
                pp[3] = __int_as_float(below2); pp[4] = __int_as_float(inside2); pp[5] = __int_as_float(anyovf2 ? 1 : 0);
            }
            grp_sync<G>(q);
            zz = 0.0f; es = 0.0f; ew = 0.0f; below2 = 0; inside2 = 0; anyovf2 = false;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float* pp = S.part[g];
                zz += pp[0]; es += pp[1]; ew += pp[2];
                below2 += __float_as_int(pp[3]); inside2 += __float_as_int(pp[4]); anyovf2 |= __float_as_int(pp[5]) != 0;
            }
        } else {
            __syncwarp();
        }
        const int nc = S.ncand;
        if (anyovf2 || nc > kGrpCand) { redo = true; why = 1; }
        if (!redo && !(r1 >= below2 && r2 < below2 + inside2)) { redo = true; why = 2; }
        redo = __any_sync(kFull, redo);

        // ---- rare extra pass over the on-chip segment: near-best count, exact variance, exp sum without the ones
        const float near_thr = near_threshold(mn, a.torch_mode);
        const double dmean = mean - (double)mn;
        const double var_z = (double)zz * inv_n_d - dmean * dmean;
        const bool need_near = !(near_thr < U);
        const bool need_var = !(var_z * 20.0 > dmean * dmean);
        const bool need_exp = es < 4.0f;
        const bool extra = __any_sync(kFull, !redo && (need_near || need_var || need_exp));
        int nnear = 0, ones = 0;
        float tss = 0.0f, small = 0.0f;
        if (extra) {
            if (__all_sync(kFull, !need_var && !need_exp)) {
                // the common case (ties at the row minimum, e.g. clamped-at-zero rows): only the near-best count
#pragma unroll 4
                for (int g = 0; g < NG; ++g) {
                    const float4 c4 = ld_again(g * 32 + lane);
                    nnear += (c4.x <= near_thr) + (c4.y <= near_thr) + (c4.z <= near_thr) + (c4.w <= near_thr);
                }
            } else {
#pragma unroll 2
                for (int g = 0; g < NG; ++g) {
                    const float4 c4 = ld_again(g * 32 + lane);
                    const float c[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float x = c[e];
                        nnear += (x <= near_thr);
                        const float dl = x - mean_f;
                        tss = fmaf(dl, dl, tss);
                        const float z = x - mn;
                        if (z > 0.0f) small += exp2_neg_fast(z * kNegLog2e); else ++ones;
                    }
                }
            }
            nnear = warp_add_i(nnear); ones = warp_add_i(ones);
            tss = warp_sum_f(tss); small = warp_sum_f(small);
        }
        // the segment is dead now: request the next row's
        if (rnext < total_rows) request(rnext);
        if (redo) {
            if (w == 0 && lane == 0) { a.redo_list[atomicAdd(a.redo_count, 1)] = (int)r; atomicAdd(a.redo_count + why, 1); }
            if (G > 1) grp_sync<G>(q);    // part[] / cand[] of this parity are not touched again before everybody is here
            continue;
        }
        if (G > 1 && extra) {
            grp_sync<G>(q);              // pass-2 partials read by everybody
            if (lane == 0) {
                float* pp = S.part[w];
                pp[6] = __int_as_float(nnear); pp[7] = __int_as_float(ones); pp[8] = tss; pp[9] = small;
            }
            grp_sync<G>(q);
            nnear = 0; ones = 0; tss = 0.0f; small = 0.0f;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float* pp = S.part[g];
                nnear += __float_as_int(pp[6]); ones += __float_as_int(pp[7]); tss += pp[8]; small += pp[9];
            }
        }

        // ---- MAD
        float da = L2, db2 = L2;
        if (!__all_sync(kFull, tie2)) {
            if (!grp_select<G, BINS>(S, bh, list, cur, q, w, L2, bracket_upper(L2, H2 - L2), r1 - below2, r2 - below2, da, db2)) {
                if (w == 0 && lane == 0) { a.redo_list[atomicAdd(a.redo_count, 1)] = (int)r; atomicAdd(a.redo_count + 3, 1); }
                continue;
            }
        }
        if (w != 0) continue;            // the finish is warp 0's; the scratch it reads has the other parity next row

        // ---- the ksel smallest entries (candidates strictly below U, then copies of U), and the finish
        int nearc = 0;
        {
            float cv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) cv[t] = lane * 4 + t < nc ? S.cand[lane * 4 + t] : INFINITY;
            nearc = (cv[0] <= near_thr) + (cv[1] <= near_thr) + (cv[2] <= near_thr) + (cv[3] <= near_thr);
            warp_bitonic_sort<4>(cv);
            // rank t lives in lane t / 4, register t % 4
            if (lane < 8) {
#pragma unroll
                for (int t = 0; t < 4; ++t) S.sorted[4 * lane + t] = 4 * lane + t < nc ? cv[t] : U;
            }
            nearc = warp_add_i(nearc);
        }
        __syncwarp();
        float* f = a.feat + ((size_t)b * n + row) * kFeatDim;
        if (lane == 0) {
            const int near = need_near ? nnear : nearc;
            double mad = ((double)da + (double)db2) * 0.5;
            if (mad < 1e-9) mad = 1e-9;
            const float gap = S.sorted[1] - S.sorted[0];
            const float range = mx - mn;                                  // exact difference of two binary32 values, rounded once
            const float comp = __fdividef(gap, range + 1e-9f);
            const float diffi = __fdividef(1.0f, range * (1.0f / (float)(n - 1)) + 1e-9f);
            double km = 0.0;
            for (int t = 0; t < 10; ++t) km += (double)S.sorted[t];
            km *= 0.1;
            double kv = 0.0;
            for (int t = 0; t < 10; ++t) { const double d = (double)S.sorted[t] - km; kv += d * d; }
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
        if (a.topv && lane < a.topk) a.topv[((size_t)b * n + row) * (size_t)a.topk + lane] = lane < ksel ? S.sorted[lane] : INFINITY;
        __syncwarp();
    }
}

}  // namespace b200lap
