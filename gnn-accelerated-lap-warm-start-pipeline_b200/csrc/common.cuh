// common.cuh -- shared device helpers for the sm_100a kernels of libb200lap.
//
// Conventions used throughout csrc/:
//   * a cost matrix lives in HBM row-major with leading dimension `ld` (elements) and
//     storage type CT (float when the caller's binary64 matrix is exactly representable
//     in binary32 -- the benchmark inputs are -- else double).  All solver arithmetic
//     widens to binary64 first, so both storage types hold the same real numbers;
//   * the library is compiled with -fmad=false: the reference solver is built without
//     FMA contraction and bit-exact duals need the same roundings;
//   * "row-resident" kernels stage one matrix row in shared memory with the bulk
//     async-copy engine (cp.async.bulk + mbarrier, SASS UBLKCP) and then make every
//     pass over it on chip.
#pragma once
#ifndef B200LAP_EMUL
#include <cuda_runtime.h>
// Kernel launches go through one macro so that the test-only SIMT interpreter
// (tests/emul/cuda_emul.h, never part of the product build) can run the same sources.
#define B200LAP_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define B200LAP_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
#include <stdint.h>
#include <math.h>

#define B200LAP_LARGE 1000000.0   /* LAP/_lapjv_cpp/lapjv.h:4 */

namespace b200lap {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// ---- widening loads ---------------------------------------------------------------------
template <typename CT> __device__ __forceinline__ double widen(CT x) { return (double)x; }

// ---- warp reductions --------------------------------------------------------------------
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(kFull, v, m); }

__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { double t = shfl_xor_d(v, o); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { double t = shfl_xor_d(v, o); v = t > v ? t : v; }
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(kFull, v); }
__device__ __forceinline__ int warp_min_i(int v) { return __reduce_min_sync(kFull, v); }
__device__ __forceinline__ int warp_max_i(int v) { return __reduce_max_sync(kFull, v); }

// Block-wide reductions with ONE barrier: every warp reduces, lane 0 publishes to a
// parity-selected slot array, and after the barrier every warp re-reduces the (<=32)
// published partials redundantly, so the result is uniform without a second barrier.
// `slots` must hold 2*32 entries; callers alternate `parity` between consecutive uses.
struct BlockRed {
    double d[2][32];
    int i[2][32];
};

__device__ __forceinline__ double block_min_d(BlockRed& r, int parity, double v) {
    v = warp_min_d(v);
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) r.d[parity][warp_id()] = v;
    __syncthreads();
    double t = lane_id() < nw ? r.d[parity][lane_id()] : INFINITY;
    return warp_min_d(t);
}
__device__ __forceinline__ double block_max_d(BlockRed& r, int parity, double v) {
    v = warp_max_d(v);
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) r.d[parity][warp_id()] = v;
    __syncthreads();
    double t = lane_id() < nw ? r.d[parity][lane_id()] : -INFINITY;
    return warp_max_d(t);
}
__device__ __forceinline__ double block_sum_d(BlockRed& r, int parity, double v) {
    v = warp_sum_d(v);
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) r.d[parity][warp_id()] = v;
    __syncthreads();
    double t = lane_id() < nw ? r.d[parity][lane_id()] : 0.0;
    return warp_sum_d(t);
}
__device__ __forceinline__ int block_sum_i(BlockRed& r, int parity, int v) {
    v = warp_sum_i(v);
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) r.i[parity][warp_id()] = v;
    __syncthreads();
    int t = lane_id() < nw ? r.i[parity][lane_id()] : 0;
    return warp_sum_i(t);
}
__device__ __forceinline__ int block_min_i(BlockRed& r, int parity, int v) {
    v = warp_min_i(v);
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) r.i[parity][warp_id()] = v;
    __syncthreads();
    int t = lane_id() < nw ? r.i[parity][lane_id()] : 0x7fffffff;
    return warp_min_i(t);
}
// min of a double together with the sum of an int, one barrier
__device__ __forceinline__ double block_min_d_sum_i(BlockRed& r, int parity, double v, int c, int* csum) {
    v = warp_min_d(v);
    c = warp_sum_i(c);
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) { r.d[parity][warp_id()] = v; r.i[parity][warp_id()] = c; }
    __syncthreads();
    double t = lane_id() < nw ? r.d[parity][lane_id()] : INFINITY;
    int ci = lane_id() < nw ? r.i[parity][lane_id()] : 0;
    *csum = warp_sum_i(ci);
    return warp_min_d(t);
}


// ---- lexicographic (value, index) top-2, the reduction behind every "two smallest" row scan --
// Entries are distinct (value, index) pairs; an empty slot is (+inf, INT_MAX).
struct Top2 {
    double a1, a2;
    int i1, i2;
};
__device__ __forceinline__ bool lex_less(double a, int ia, double b, int ib) { return a < b || (a == b && ia < ib); }
__device__ __forceinline__ void top2_init(Top2& t) { t.a1 = INFINITY; t.a2 = INFINITY; t.i1 = 0x7fffffff; t.i2 = 0x7fffffff; }
__device__ __forceinline__ void top2_push(Top2& t, double a, int i) {
    if (lex_less(a, i, t.a1, t.i1)) { t.a2 = t.a1; t.i2 = t.i1; t.a1 = a; t.i1 = i; }
    else if (lex_less(a, i, t.a2, t.i2)) { t.a2 = a; t.i2 = i; }
}
__device__ __forceinline__ void top2_merge(Top2& t, const Top2& o) {
    top2_push(t, o.a1, o.i1);
    top2_push(t, o.a2, o.i2);
}
// for keys pushed in increasing index order (a thread walking its own columns): ties keep the earlier index by themselves
__device__ __forceinline__ void top2_push_inc(Top2& t, double a, int i) {
    if (a < t.a1) { t.a2 = t.a1; t.i2 = t.i1; t.a1 = a; t.i1 = i; }
    else if (a < t.a2) { t.a2 = a; t.i2 = i; }
}
// order-preserving 64-bit image of a binary64 value (-0 folded onto +0 first: they compare equal)
__device__ __forceinline__ unsigned long long d2ord(double x) {
    const long long b = __double_as_longlong(x + 0.0);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
// lane holding the lexicographic minimum of (a, i) over the warp: three integer redux steps instead of a
// five-round shuffle butterfly of 64-bit values
__device__ __forceinline__ int warp_lexmin_lane(double a, int i) {
    const unsigned long long k = d2ord(a);
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_min_sync(kFull, hi);
    const bool c1 = hi == mh;
    const unsigned ml = __reduce_min_sync(kFull, c1 ? lo : 0xffffffffu);
    const bool c2 = c1 && lo == ml;
    const unsigned mi = __reduce_min_sync(kFull, c2 ? (unsigned)i : 0xffffffffu);
    const unsigned who = __ballot_sync(kFull, c2 && (unsigned)i == mi);
    return __ffs((int)who) - 1;
}
__device__ __forceinline__ Top2 warp_top2(Top2 t) {
    Top2 r;
    const int w1 = warp_lexmin_lane(t.a1, t.i1);
    r.a1 = __shfl_sync(kFull, t.a1, w1);
    r.i1 = __shfl_sync(kFull, t.i1, w1);
    if (lane_id() == w1) { t.a1 = t.a2; t.i1 = t.i2; }     // the runner-up is the winner's second or another lane's first
    const int w2 = warp_lexmin_lane(t.a1, t.i1);
    r.a2 = __shfl_sync(kFull, t.a1, w2);
    r.i2 = __shfl_sync(kFull, t.i1, w2);
    return r;
}

// ---- a block-reduction scratch with the slot parity kept alongside ----------------------------
struct BlockRed2 {
    double d2[2][32];
    int i2[2][32];
    double out_a[2][2];   // block_top2 result (a1, a2) per parity
    int out_i[2][2];
};
struct Red {
    BlockRed* r;
    BlockRed2* r2;
    int par;
    __device__ __forceinline__ int flip() { par ^= 1; return par; }
};
__device__ __forceinline__ Top2 block_top2(Red& R, Top2 t) {
    t = warp_top2(t);
    const int p = R.flip();
    const int nw = (blockDim.x + 31) >> 5;
    if (lane_id() == 0) {
        R.r->d[p][warp_id()] = t.a1; R.r->i[p][warp_id()] = t.i1;
        R.r2->d2[p][warp_id()] = t.a2; R.r2->i2[p][warp_id()] = t.i2;
    }
    __syncthreads();
    if (nw == 1) return t;
    // the cross-warp step is done once, by warp 0, and broadcast (it used to be repeated by every warp)
    if (warp_id() == 0) {
        Top2 q;
        top2_init(q);
        if (lane_id() < nw) {
            q.a1 = R.r->d[p][lane_id()]; q.i1 = R.r->i[p][lane_id()];
            q.a2 = R.r2->d2[p][lane_id()]; q.i2 = R.r2->i2[p][lane_id()];
        }
        // first of the union = lexmin of the firsts; second = lexmin of (winner's second, the other firsts)
        const int w1 = warp_lexmin_lane(q.a1, q.i1);
        const double a1 = __shfl_sync(kFull, q.a1, w1);
        const int i1 = __shfl_sync(kFull, q.i1, w1);
        if (lane_id() == w1) { q.a1 = q.a2; q.i1 = q.i2; }
        const int w2 = warp_lexmin_lane(q.a1, q.i1);
        const double a2 = __shfl_sync(kFull, q.a1, w2);
        const int i2 = __shfl_sync(kFull, q.i1, w2);
        if (lane_id() == 0) { R.r2->out_a[p][0] = a1; R.r2->out_a[p][1] = a2; R.r2->out_i[p][0] = i1; R.r2->out_i[p][1] = i2; }
    }
    __syncthreads();
    Top2 r;
    r.a1 = R.r2->out_a[p][0]; r.a2 = R.r2->out_a[p][1]; r.i1 = R.r2->out_i[p][0]; r.i2 = R.r2->out_i[p][1];
    return r;
}
__device__ __forceinline__ double red_min_d(Red& R, double v) { return block_min_d(*R.r, R.flip(), v); }
__device__ __forceinline__ double red_max_d(Red& R, double v) { return block_max_d(*R.r, R.flip(), v); }
__device__ __forceinline__ double red_sum_d(Red& R, double v) { return block_sum_d(*R.r, R.flip(), v); }
__device__ __forceinline__ int red_sum_i(Red& R, int v) { return block_sum_i(*R.r, R.flip(), v); }
__device__ __forceinline__ int red_min_i(Red& R, int v) { return block_min_i(*R.r, R.flip(), v); }
__device__ __forceinline__ double red_min_d_sum_i(Red& R, double v, int c, int* cs) { return block_min_d_sum_i(*R.r, R.flip(), v, c, cs); }

#ifndef B200LAP_EMUL
// ---- mbarrier + bulk async copy (global -> shared), the 1-D TMA path ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// bytes must be a multiple of 16, src/dst 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// order prior generic-proxy accesses of shared memory before later async-proxy writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#endif  // !B200LAP_EMUL

// ---- misc ---------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

}  // namespace b200lap
