// mlp_tc.cuh -- OneGNN forward with the hidden-192 GEMMs on the 5th-gen tensor cores (tcgen05).
//
// Reference: gnn/one_gnn.py:89-160 (same math as mlp.cuh, which remains the path for other widths).
//
// A CTA owns 128 rows (= the 128 TMEM lanes).  Every layer  D[128 x N] = A[128 x K] . W^T  runs as
// tcgen05.mma.kind::tf32 with the accumulator in tensor memory.  binary32-level accuracy comes from
// the 3xTF32 split  a = a_hi + a_lo, w = w_hi + w_lo  (hi = top 19 bits, lo = exact remainder):
//     D = a_hi w_hi + a_hi w_lo + a_lo w_hi          (the dropped a_lo w_lo term is ~2^-22 relative)
//   * a_hi lives in shared memory in the UMMA K-major no-swizzle layout  [k/4][row][4]  (96 KB),
//   * a_lo lives in TENSOR MEMORY (columns 256..447) and is fed as the A operand from TMEM,
//   * w_hi/w_lo are pre-split and pre-tiled at model creation; 16-wide K chunks (24 KB hi+lo) stream
//     through a 4-stage shared ring with one cp.async.bulk (TMA engine) + mbarrier per chunk.
// Warp roles: warp 8 = control (lane 0 issues the bulk copies and the MMAs, commits to mbarriers);
// warps 0-7 = epilogue: warp w reads TMEM lanes 32(w%4).. (its rows) and columns 96(w/4).. with
// tcgen05.ld, applies bias / GELU / LayerNorm / residual in registers, splits the result and writes
// the next layer's a_hi (st.shared) and a_lo (tcgen05.st).  The residual stream stays in registers.
#pragma once
#include "common.cuh"
#include "mlp.cuh"

namespace b200lap {

constexpr int kTcRows = 128;
constexpr int kTcH = 192;
constexpr int kTcEpiThreads = 256;
constexpr int kTcThreads = kTcEpiThreads + 32;
constexpr int kTcKC = 16;
constexpr int kTcStages = 4;
constexpr int kTcSlabA = kTcRows * 16;                    // bytes of one k/4 slab of A
constexpr int kTcStageBytes = 2 * (kTcKC / 4) * kTcH * 16;   // hi + lo chunk at N = 192
constexpr int kTcMaxGemm = 12;
constexpr int kTcAloCol = 256;                            // TMEM column where a_lo starts
constexpr int kTcTmemCols = 512;

struct TcGemms {
    const float* w[kTcMaxGemm];   // pre-tiled hi/lo chunks
    int K[kTcMaxGemm];            // padded to a multiple of kTcKC
    int N[kTcMaxGemm];
    int edge_index;               // index of the edge-MLP GEMM (skipped when the cost branch is off)
    int count;                    // number of GEMMs including the edge one
};

__host__ __device__ inline size_t tc_smem_bytes() {
    return (size_t)(kTcH / 4) * kTcSlabA + (size_t)kTcStages * kTcStageBytes + 4096;
}

// floats of one pre-tiled GEMM: K/16 chunks x (hi + lo) x 4 slabs x N x 4
__host__ __device__ inline size_t tc_gemm_floats(int K, int N) { return (size_t)(K / kTcKC) * 2 * (kTcKC / 4) * N * 4; }

// W [N][K_in] (torch layout) -> chunked hi/lo tiles (model creation only)
__global__ void k_tc_pack_weights(const float* __restrict__ W, int N, int K_in, int K_pad, float* __restrict__ dst)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = K_pad * N;
    if (idx >= total) return;
    const int k = idx / N, n = idx % N;
    const float v = k < K_in ? W[(size_t)n * K_in + k] : 0.0f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const float lo = v - hi;
    const int c = k / kTcKC, s = (k % kTcKC) / 4, q = k % 4;
    const size_t chunk = (size_t)c * 2 * (kTcKC / 4) * N * 4;
    const size_t off = ((size_t)s * N + n) * 4 + q;
    dst[chunk + off] = hi;
    dst[chunk + (size_t)(kTcKC / 4) * N * 4 + off] = lo;
}

#ifndef B200LAP_EMUL
// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout SWIZZLE_NONE [61,64)
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t tc_idesc_tf32(int M, int N) {
    // cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1 | a_format TF32 [7,10)=2 | b_format [10,13)=2 |
    // a/b K-major | n_dim = N>>3 [17,23) | m_dim = M>>4 [24,29)
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory"); }

struct TcShared {
    uint64_t full[kTcStages];
    uint64_t empty[kTcStages];
    uint64_t acc_full;
    uint64_t act_ready;
    uint32_t tmem_base;
    float part[2][2][kTcRows];   // [buffer][column half][row]: partial row sums exchanged between the two halves
};

// ---- epilogue helpers (one thread = one row x 96 columns) -------------------------------------------
struct EpiCtx {
    TcShared* sh;
    unsigned char* a_hi;    // shared A tile
    uint32_t tmem;          // TMEM base (lane 0, column 0)
    int row, half, quad;    // row in tile, column half, TMEM lane quadrant
    int xbuf;               // exchange buffer toggle
};

// sum of a per-thread partial over the two column halves of the row
__device__ __forceinline__ float row_sum2(EpiCtx& E, float v) {
    E.sh->part[E.xbuf][E.half][E.row] = v;
    epi_bar();
    const float o = E.sh->part[E.xbuf][E.half ^ 1][E.row];
    E.xbuf ^= 1;
    return v + o;
}

__device__ __forceinline__ void load_acc96(const EpiCtx& E, float (&acc)[96]) {
    const uint32_t t = E.tmem + ((uint32_t)(E.quad * 32) << 16) + (uint32_t)(E.half * 96);
    float a0[32], a1[32], a2[32];
    tc_ld32(t, a0);
    tc_ld32(t + 32, a1);
    tc_ld32(t + 64, a2);
    tc_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) { acc[i] = a0[i]; acc[32 + i] = a1[i]; acc[64 + i] = a2[i]; }
}

// split v into hi (shared A tile) and lo (TMEM), columns [col0, col0 + CNT) of the next layer's K
template <int CNT>
__device__ __forceinline__ void store_act(const EpiCtx& E, const float (&v)[CNT], int col0) {
    static_assert(CNT % 32 == 0, "whole TMEM stores");
#pragma unroll
    for (int g = 0; g < CNT / 32; ++g) {
        float lo[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float h4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float x = v[g * 32 + i + q];
                h4[q] = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
                lo[i + q] = x - h4[q];
            }
            const int k4 = (col0 + g * 32 + i) >> 2;
            *reinterpret_cast<float4*>(E.a_hi + (size_t)k4 * kTcSlabA + (size_t)E.row * 16) = make_float4(h4[0], h4[1], h4[2], h4[3]);
        }
        tc_st32(E.tmem + ((uint32_t)(E.quad * 32) << 16) + (uint32_t)(kTcAloCol + col0 + g * 32), lo);
    }
}

// publish the freshly written activations to the MMA (async proxy) and signal the control warp
__device__ __forceinline__ void publish_act(const EpiCtx& E) {
    tc_st_wait();
    tc_fence_before();
    fence_proxy_async();
    mbar_arrive(&E.sh->act_ready);
}

__device__ __forceinline__ void layer_norm96(EpiCtx& E, float (&x)[96], const float* __restrict__ g, const float* __restrict__ b) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 96; ++i) s += x[i];
    const float mean = row_sum2(E, s) * (1.0f / (float)kTcH);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 96; ++i) { const float t = x[i] - mean; q = fmaf(t, t, q); }
    const float var = row_sum2(E, q) * (1.0f / (float)kTcH);
    const float rs = 1.0f / sqrtf(var + 1e-5f);
    const int c0 = E.half * 96;
#pragma unroll
    for (int i = 0; i < 96; ++i) x[i] = fmaf((x[i] - mean) * rs, __ldg(g + c0 + i), __ldg(b + c0 + i));
}

__global__ void __launch_bounds__(kTcThreads, 1) k_onegnn_tc(
    MlpWeights W, TcGemms G, const float* __restrict__ feat /* [B][n][in_dim] */, const float* __restrict__ topv /* [B][n][topk] */,
    int has_cost, int n, float* __restrict__ raw /* [B][n] */)
{
    B200LAP_DYN_SMEM(dyn);
    unsigned char* a_hi = dyn;                                        // 48 slabs x 2 KB
    unsigned char* stages = dyn + (size_t)(kTcH / 4) * kTcSlabA;      // 4 x 24 KB
    TcShared* sh = reinterpret_cast<TcShared*>(stages + (size_t)kTcStages * kTcStageBytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, row0 = blockIdx.x * kTcRows;

    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
        mbar_init(&sh->acc_full, 1);
        mbar_init(&sh->act_ready, kTcEpiThreads);
        mbar_fence_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"((uint32_t)kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh->tmem_base;
    const int n_gemm = has_cost ? G.count : G.count - 1;

    if (warp == 8) {
        // ================= control warp: weight stream + MMA issue =================
        if (lane == 0) {
            int total = 0;
            for (int g = 0; g < G.count; ++g)
                if (has_cost || g != G.edge_index) total += G.K[g] / kTcKC;
            // load cursor
            int lg = 0, lc = 0, t_load = 0;
            auto skip_load = [&]() { if (!has_cost && lg == G.edge_index) { ++lg; lc = 0; } };
            auto issue_load = [&]() {
                skip_load();
                const int s = t_load % kTcStages;
                const int N = G.N[lg];
                const uint32_t bytes = (uint32_t)(2 * (kTcKC / 4) * N * 16);
                mbar_expect_tx(&sh->full[s], bytes);
                bulk_g2s(stages + (size_t)s * kTcStageBytes, G.w[lg] + (size_t)lc * (bytes / 4), bytes, &sh->full[s]);
                ++t_load;
                if (++lc == G.K[lg] / kTcKC) { ++lg; lc = 0; }
            };
            for (int i = 0; i < kTcStages && t_load < total; ++i) issue_load();
            int t = 0, layer = 0;
            for (int g = 0; g < G.count; ++g) {
                if (!has_cost && g == G.edge_index) continue;
                const int N = G.N[g], nchunk = G.K[g] / kTcKC;
                const uint32_t idesc = tc_idesc_tf32(kTcRows, N);
                mbar_wait(&sh->act_ready, (unsigned)(layer & 1));      // the epilogue wrote this layer's input
                tc_fence_after();
                for (int c = 0; c < nchunk; ++c, ++t) {
                    const int s = t % kTcStages;
                    mbar_wait(&sh->full[s], (unsigned)((t / kTcStages) & 1));
                    tc_fence_after();
                    const uint32_t sb = smem_u32(stages + (size_t)s * kTcStageBytes);
                    const uint32_t lbo_b = (uint32_t)N * 16;
#pragma unroll
                    for (int ks = 0; ks < kTcKC / 8; ++ks) {
                        const int k0 = c * kTcKC + ks * 8;
                        const uint64_t ad = tc_smem_desc(smem_u32(a_hi) + (uint32_t)(k0 / 4) * kTcSlabA, kTcSlabA, 128);
                        const uint64_t bhi = tc_smem_desc(sb + (uint32_t)(ks * 2) * lbo_b, lbo_b, 128);
                        const uint64_t blo = tc_smem_desc(sb + (uint32_t)(kTcKC / 4) * lbo_b + (uint32_t)(ks * 2) * lbo_b, lbo_b, 128);
                        const uint32_t first = (c == 0 && ks == 0) ? 0u : 1u;
                        tc_mma_ss(tmem, ad, blo, idesc, first);
                        tc_mma_ts(tmem, tmem + (uint32_t)(kTcAloCol + k0), bhi, idesc, 1u);
                        tc_mma_ss(tmem, ad, bhi, idesc, 1u);
                    }
                    tc_commit(&sh->empty[s]);
                    if (c == nchunk - 1) tc_commit(&sh->acc_full);
                    if (t >= 1 && t_load < total) {
                        const int sl = t_load % kTcStages;
                        mbar_wait(&sh->empty[sl], (unsigned)(((t_load / kTcStages) - 1) & 1));
                        issue_load();
                    }
                }
                ++layer;
            }
        }
    } else {
        // ================= epilogue warps =================
        EpiCtx E;
        E.sh = sh; E.a_hi = a_hi; E.tmem = tmem;
        E.quad = warp & 3; E.half = warp >> 2; E.row = E.quad * 32 + lane; E.xbuf = 0;
        const int grow = row0 + E.row;
        const int c0 = E.half * 96;
        int layer = 0;
        // ---- stage the feature tile (K padded to 32: half 0 -> columns 0..15, half 1 -> 16..31)
        {
            float x[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = 0.0f;
            if (grow < n) {
                const float* fr = feat + ((size_t)b * n + grow) * W.in_dim;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int k = E.half * 16 + i;
                    if (k < W.in_dim) x[i] = __ldg(fr + k);
                }
            }
            // columns [16*half, 16*half+16): write hi to shared, lo to TMEM (a 32-wide store covers both halves' ranges:
            // each half stores only its own 16 columns, padded slots are zero)
            float lo[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                float h4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    h4[q] = __uint_as_float(__float_as_uint(x[i + q]) & 0xffffe000u);
                    lo[i + q] = x[i + q] - h4[q];
                }
                const int k4 = (E.half * 16 + i) >> 2;
                *reinterpret_cast<float4*>(a_hi + (size_t)k4 * kTcSlabA + (size_t)E.row * 16) = make_float4(h4[0], h4[1], h4[2], h4[3]);
            }
            const uint32_t ta = tmem + ((uint32_t)(E.quad * 32) << 16) + (uint32_t)(kTcAloCol + E.half * 16);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta),
                "r"(__float_as_uint(lo[0])), "r"(__float_as_uint(lo[1])), "r"(__float_as_uint(lo[2])), "r"(__float_as_uint(lo[3])),
                "r"(__float_as_uint(lo[4])), "r"(__float_as_uint(lo[5])), "r"(__float_as_uint(lo[6])), "r"(__float_as_uint(lo[7])),
                "r"(__float_as_uint(lo[8])), "r"(__float_as_uint(lo[9])), "r"(__float_as_uint(lo[10])), "r"(__float_as_uint(lo[11])),
                "r"(__float_as_uint(lo[12])), "r"(__float_as_uint(lo[13])), "r"(__float_as_uint(lo[14])), "r"(__float_as_uint(lo[15]))
                : "memory");
            publish_act(E);
        }
        float h[96], acc[96];
        auto wait_acc = [&]() {
            mbar_wait(&sh->acc_full, (unsigned)(layer & 1));
            tc_fence_after();
            ++layer;
        };
        // ---- input projection: h = LN(GELU(x W_in^T + b))
        wait_acc();
        load_acc96(E, acc);
#pragma unroll
        for (int i = 0; i < 96; ++i) h[i] = gelu_erf(acc[i] + __ldg(W.b_in + c0 + i));
        layer_norm96(E, h, W.ln0_g, W.ln0_b);
        store_act<96>(E, h, c0);
        publish_act(E);
        // ---- residual blocks
        for (int l = 0; l < W.layers; ++l) {
            const float* p = W.blk + (size_t)l * mlp_block_floats(kTcH);
            const float* b1 = p + (size_t)kTcH * kTcH;
            const float* b2 = b1 + kTcH + (size_t)kTcH * kTcH;
            const float* g = b2 + kTcH;
            const float* be = g + kTcH;
            wait_acc();
            load_acc96(E, acc);
#pragma unroll
            for (int i = 0; i < 96; ++i) acc[i] = gelu_erf(acc[i] + __ldg(b1 + c0 + i));
            store_act<96>(E, acc, c0);
            publish_act(E);
            wait_acc();
            load_acc96(E, acc);
#pragma unroll
            for (int i = 0; i < 96; ++i) h[i] = h[i] + (acc[i] + __ldg(b2 + c0 + i));
            layer_norm96(E, h, g, be);
            store_act<96>(E, h, c0);
            if (l + 1 < W.layers || !has_cost) publish_act(E);   // with the cost branch the edge input replaces it below
        }
        if (has_cost) {
            // ---- u_pre and the sparse top-k refinement (gnn/one_gnn.py:122-160)
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < 96; ++i) s = fmaf(h[i], __ldg(W.pre_w + c0 + i), s);
            const float upre = row_sum2(E, s) + W.pre_b;
#pragma unroll
            for (int i = 0; i < 96; ++i) acc[i] = 0.0f;
            float wsum = 0.0f;
            if (grow < n) {
                const float* tv = topv + ((size_t)b * n + grow) * W.topk;
                float mx = -INFINITY;
                for (int k = 0; k < W.topk; ++k) {
                    const float val = __ldg(tv + k) - upre;
                    if (isfinite(val)) mx = fmaxf(mx, -val);
                }
                if (!isfinite(mx)) mx = 0.0f;
                float den = 0.0f;
                for (int k = 0; k < W.topk; ++k) {
                    const float val = __ldg(tv + k) - upre;
                    if (isfinite(val)) den += expf(-val - mx);
                }
                for (int k = 0; k < W.topk; ++k) {
                    const float val = __ldg(tv + k) - upre;
                    if (!isfinite(val)) continue;
                    const float wk = expf(-val - mx) / den;
                    wsum += wk;
#pragma unroll
                    for (int i = 0; i < 96; ++i)
                        acc[i] = fmaf(wk, gelu_erf(fmaf(__ldg(W.e1_w + c0 + i), val, __ldg(W.e1_b + c0 + i))), acc[i]);
                }
            }
            store_act<96>(E, acc, c0);
            publish_act(E);
            wait_acc();
            load_acc96(E, acc);
#pragma unroll
            for (int i = 0; i < 96; ++i) acc[i] = fmaf(__ldg(W.e2_b + c0 + i), wsum, acc[i]);
            layer_norm96(E, acc, W.mg, W.mb);
#pragma unroll
            for (int i = 0; i < 96; ++i) h[i] += acc[i];
            store_act<96>(E, h, c0);
            publish_act(E);
        }
        // ---- head: Linear(192, 96) -> GELU -> Linear(96, 1); this half owns output columns [48*half, +48)
        wait_acc();
        {
            const uint32_t t = tmem + ((uint32_t)(E.quad * 32) << 16) + (uint32_t)(E.half * 48);
            float a0[32], a1[16];
            tc_ld32(t, a0);
            tc_ld16(t + 32, a1);
            tc_ld_wait();
            float s = 0.0f;
            const int h0 = E.half * 48;
#pragma unroll
            for (int i = 0; i < 32; ++i) s = fmaf(gelu_erf(a0[i] + __ldg(W.ra_b + h0 + i)), __ldg(W.rb_w + h0 + i), s);
#pragma unroll
            for (int i = 0; i < 16; ++i) s = fmaf(gelu_erf(a1[i] + __ldg(W.ra_b + h0 + 32 + i)), __ldg(W.rb_w + h0 + 32 + i), s);
            s = row_sum2(E, s) + W.rb_b;
            if (E.half == 0 && grow < n) raw[(size_t)b * n + grow] = s;
        }
        (void)n_gemm;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kTcTmemCols) : "memory");
    }
}
#endif  // !B200LAP_EMUL

}  // namespace b200lap
