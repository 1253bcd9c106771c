// mlp.cuh -- OneGNN inference forward on row tiles (binary32, FFMA path).
//
// Reference: gnn/one_gnn.py:89-160 (OneGNN.forward + _sparse_refine), :18-36 (ResidualBlock).
//   h  = LN(GELU(x W_in^T + b))                       input_proj
//   h  = LN(h + fc2(GELU(fc1 h)))   x layers          blocks
//   u_pre = pre_out(h)
//   vals = fl32(top-k smallest cost - u_pre); w = softmax(-vals)
//   msg = sum_k w_k edge_mlp(vals_k) = E2 (sum_k w_k GELU(e1 vals_k + b1)) + b2 sum_k w_k   (linear second layer
//         applied after the k-sum: 16x fewer FLOPs, rounding-level difference -- SURVEY.md App. C.2)
//   h  = h + LN(msg);  raw = row_out(h);  u = raw - mean(raw)
//
// A CTA owns a tile of kTileRows rows; activations stay in shared memory k-major ([feature][row])
// so a thread's 4 rows are one 128-bit load, weights stream through a double-buffered shared
// chunk ([k][out], pre-transposed at model creation), the residual stream h lives in registers.
// 256 threads = 16 (row groups) x 16 (column groups); a row's LayerNorm/dot reductions are
// shuffles over the 16 lanes that share the row group.
#pragma once
#include "common.cuh"

namespace b200lap {

constexpr int kTileRows = 64;
constexpr int kMlpThreads = 256;
constexpr int kKC = 16;        // weight rows per shared chunk
constexpr int kInPad = 32;     // input features padded to a whole number of chunks

struct MlpWeights {
    int in_dim, hidden, layers, topk;
    const float* w_in_t;   // [kInPad][H]
    const float* b_in;     // [H]
    const float* ln0_g; const float* ln0_b;
    const float* blk;      // per block: W1t[H][H], b1[H], W2t[H][H], b2[H], g[H], b[H]
    const float* pre_w;    // [H]
    float pre_b;
    const float* ra_t;     // [H][H/2]
    const float* ra_b;     // [H/2]
    const float* rb_w;     // [H/2]
    float rb_b;
    const float* e1_w; const float* e1_b;   // [H]
    const float* e2_t;     // [H][H]
    const float* e2_b;     // [H]
    const float* mg; const float* mb;       // [H]
};

__host__ __device__ inline size_t mlp_block_floats(int H) { return (size_t)2 * H * H + (size_t)4 * H; }
__host__ __device__ inline size_t mlp_smem_bytes(int H) { return ((size_t)H * kTileRows + (size_t)2 * kKC * H) * sizeof(float); }

template <int CP> __device__ __forceinline__ void load_cols(const float* p, float (&w)[CP]) {
    if constexpr (CP % 4 == 0) {
#pragma unroll
        for (int q = 0; q < CP / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
    } else if constexpr (CP % 2 == 0) {
#pragma unroll
        for (int q = 0; q < CP / 2; ++q) {
            const float2 t = *reinterpret_cast<const float2*>(p + 2 * q);
            w[2 * q] = t.x; w[2 * q + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < CP; ++q) w[q] = p[q];
    }
}

// sum over the 16 lanes that share a row group
__device__ __forceinline__ float group16_sum(float v) {
    v += __shfl_xor_sync(kFull, v, 8);
    v += __shfl_xor_sync(kFull, v, 4);
    v += __shfl_xor_sync(kFull, v, 2);
    v += __shfl_xor_sync(kFull, v, 1);
    return v;
}

// acc[4][CP] = act[K][tile rows of this thread] x Wt[K][N] (columns tx*CP ..)
template <int N, int CP>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ act, const float* __restrict__ Wt, int K, float* wbuf,
                                          float (&acc)[4][CP])
{
    static_assert(N == CP * 16, "column ownership");
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    constexpr int kChunkFloats = kKC * N;
    constexpr int kPerThread = (kChunkFloats + kMlpThreads - 1) / kMlpThreads;
    float stage[kPerThread];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[r][c] = 0.0f;
    const int nchunks = K / kKC;
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
        const int idx = q * kMlpThreads + tid;
        if (idx < kChunkFloats) wbuf[idx] = __ldg(Wt + idx);
    }
    __syncthreads();
    for (int ch = 0; ch < nchunks; ++ch) {
        const float* wcur = wbuf + (ch & 1) * kChunkFloats;
        if (ch + 1 < nchunks) {
#pragma unroll
            for (int q = 0; q < kPerThread; ++q) {
                const int idx = q * kMlpThreads + tid;
                if (idx < kChunkFloats) stage[q] = __ldg(Wt + (size_t)(ch + 1) * kChunkFloats + idx);
            }
        }
        const float* arow = act + (size_t)ch * kKC * kTileRows + ty * 4;
#pragma unroll
        for (int kk = 0; kk < kKC; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(arow + kk * kTileRows);
            float w[CP];
            load_cols<CP>(wcur + kk * N + tx * CP, w);
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                acc[0][c] = fmaf(a.x, w[c], acc[0][c]);
                acc[1][c] = fmaf(a.y, w[c], acc[1][c]);
                acc[2][c] = fmaf(a.z, w[c], acc[2][c]);
                acc[3][c] = fmaf(a.w, w[c], acc[3][c]);
            }
        }
        if (ch + 1 < nchunks) {
            float* wnext = wbuf + ((ch + 1) & 1) * kChunkFloats;
#pragma unroll
            for (int q = 0; q < kPerThread; ++q) {
                const int idx = q * kMlpThreads + tid;
                if (idx < kChunkFloats) wnext[idx] = stage[q];
            }
        }
        __syncthreads();
    }
}

template <int CP>
__device__ __forceinline__ void store_act(float* act, const float (&v)[4][CP]) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int c = 0; c < CP; ++c)
        *reinterpret_cast<float4*>(act + (size_t)(tx * CP + c) * kTileRows + ty * 4) = make_float4(v[0][c], v[1][c], v[2][c], v[3][c]);
}

// in-place LayerNorm of the 4 rows held by the 16-lane group (biased variance, eps 1e-5)
template <int H, int CP>
__device__ __forceinline__ void layer_norm_rows(float (&v)[4][CP], const float* __restrict__ g, const float* __restrict__ b) {
    const int tx = threadIdx.x & 15;
    float gg[CP], bb[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) { gg[c] = __ldg(g + tx * CP + c); bb[c] = __ldg(b + tx * CP + c); }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < CP; ++c) s += v[r][c];
        const float mean = group16_sum(s) * (1.0f / (float)H);
        float q = 0.0f;
#pragma unroll
        for (int c = 0; c < CP; ++c) { const float t = v[r][c] - mean; q = fmaf(t, t, q); }
        const float var = group16_sum(q) * (1.0f / (float)H);
        const float rs = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
        for (int c = 0; c < CP; ++c) v[r][c] = fmaf((v[r][c] - mean) * rs, gg[c], bb[c]);
    }
}

template <int H>
__global__ void __launch_bounds__(kMlpThreads) k_onegnn_tile(
    MlpWeights W, const float* __restrict__ feat /* [B][n][in_dim] */, const float* __restrict__ topv /* [B][n][topk] */,
    int has_cost, int n, float* __restrict__ raw /* [B][n] */)
{
    constexpr int CP = H / 16;
    constexpr int H2 = H / 2;
    constexpr int CP2 = H2 / 16;
    B200LAP_DYN_SMEM(dyn);
    float* act = reinterpret_cast<float*>(dyn);
    float* wbuf = act + (size_t)H * kTileRows;
    const int b = blockIdx.y, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int row0 = blockIdx.x * kTileRows;
    const float* fb = feat + (size_t)b * n * W.in_dim;
    // ---- stage the feature tile k-major, zero padded to kInPad features / kTileRows rows
    for (int idx = tid; idx < kInPad * kTileRows; idx += kMlpThreads) {
        const int k = idx / kTileRows, r = idx % kTileRows;
        act[idx] = (k < W.in_dim && row0 + r < n) ? __ldg(fb + (size_t)(row0 + r) * W.in_dim + k) : 0.0f;
    }
    float h[4][CP], acc[4][CP];
    // ---- input projection
    tile_gemm<H, CP>(act, W.w_in_t, kInPad, wbuf, acc);
#pragma unroll
    for (int c = 0; c < CP; ++c) {
        const float bias = __ldg(W.b_in + tx * CP + c);
#pragma unroll
        for (int r = 0; r < 4; ++r) h[r][c] = gelu_erf(acc[r][c] + bias);
    }
    layer_norm_rows<H, CP>(h, W.ln0_g, W.ln0_b);
    store_act<CP>(act, h);
    // ---- residual blocks
    for (int l = 0; l < W.layers; ++l) {
        const float* p = W.blk + (size_t)l * mlp_block_floats(H);
        const float* w1 = p; const float* b1 = w1 + (size_t)H * H;
        const float* w2 = b1 + H; const float* b2 = w2 + (size_t)H * H;
        const float* g = b2 + H; const float* be = g + H;
        tile_gemm<H, CP>(act, w1, H, wbuf, acc);
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            const float bias = __ldg(b1 + tx * CP + c);
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r][c] = gelu_erf(acc[r][c] + bias);
        }
        store_act<CP>(act, acc);
        tile_gemm<H, CP>(act, w2, H, wbuf, acc);
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            const float bias = __ldg(b2 + tx * CP + c);
#pragma unroll
            for (int r = 0; r < 4; ++r) h[r][c] = h[r][c] + (acc[r][c] + bias);
        }
        layer_norm_rows<H, CP>(h, g, be);
        store_act<CP>(act, h);
    }
    if (has_cost) {
        // ---- u_pre and the sparse top-k refinement
        float upre[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float s = 0.0f;
#pragma unroll
            for (int c = 0; c < CP; ++c) s = fmaf(h[r][c], __ldg(W.pre_w + tx * CP + c), s);
            upre[r] = group16_sum(s) + W.pre_b;
        }
        float e1w[CP], e1b[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) { e1w[c] = __ldg(W.e1_w + tx * CP + c); e1b[c] = __ldg(W.e1_b + tx * CP + c); }
        float wsum[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = row0 + ty * 4 + r;
#pragma unroll
            for (int c = 0; c < CP; ++c) acc[r][c] = 0.0f;
            wsum[r] = 0.0f;
            if (row < n) {
                const float* tv = topv + ((size_t)b * n + row) * W.topk;
                float mx = -INFINITY;
                for (int k = 0; k < W.topk; ++k) {
                    const float val = __ldg(tv + k) - upre[r];
                    if (isfinite(val)) mx = fmaxf(mx, -val);
                }
                if (!isfinite(mx)) mx = 0.0f;
                float den = 0.0f;
                for (int k = 0; k < W.topk; ++k) {
                    const float val = __ldg(tv + k) - upre[r];
                    if (isfinite(val)) den += expf(-val - mx);
                }
                for (int k = 0; k < W.topk; ++k) {
                    const float val = __ldg(tv + k) - upre[r];
                    if (!isfinite(val)) continue;
                    const float wk = expf(-val - mx) / den;
                    wsum[r] += wk;
#pragma unroll
                    for (int c = 0; c < CP; ++c) acc[r][c] = fmaf(wk, gelu_erf(fmaf(e1w[c], val, e1b[c])), acc[r][c]);
                }
            }
        }
        __syncthreads();
        store_act<CP>(act, acc);
        tile_gemm<H, CP>(act, W.e2_t, H, wbuf, acc);
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            const float bias = __ldg(W.e2_b + tx * CP + c);
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r][c] = fmaf(bias, wsum[r], acc[r][c]);
        }
        layer_norm_rows<H, CP>(acc, W.mg, W.mb);
#pragma unroll
        for (int c = 0; c < CP; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) h[r][c] += acc[r][c];
        store_act<CP>(act, h);
    }
    // ---- head: Linear(H, H/2) -> GELU -> Linear(H/2, 1)
    float hacc[4][CP2];
    tile_gemm<H2, CP2>(act, W.ra_t, H, wbuf, hacc);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < CP2; ++c) {
            const float t = gelu_erf(hacc[r][c] + __ldg(W.ra_b + tx * CP2 + c));
            s = fmaf(t, __ldg(W.rb_w + tx * CP2 + c), s);
        }
        s = group16_sum(s) + W.rb_b;
        const int row = row0 + ty * 4 + r;
        if (tx == 0 && row < n) raw[(size_t)b * n + row] = s;
    }
}

// u = raw - mean(raw) per instance (gnn/one_gnn.py:112-113); also widens to binary64 for the solver
__global__ void __launch_bounds__(1024) k_center_rows(const float* __restrict__ raw, int n, float* __restrict__ u32,
                                                      double* __restrict__ u64)
{
    __shared__ BlockRed red;
    const int b = blockIdx.x;
    const float* r = raw + (size_t)b * n;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)r[i];
    s = block_sum_d(red, 0, s);
    const float mean = (float)(s / (double)n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float u = r[i] - mean;
        if (u32) u32[(size_t)b * n + i] = u;
        if (u64) u64[(size_t)b * n + i] = (double)u;
    }
}

// [out][in] -> [in_pad][out], zero padded rows (model creation only)
__global__ void k_transpose_pad(const float* __restrict__ src, int out_dim, int in_dim, int in_pad, float* __restrict__ dst)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= in_pad * out_dim) return;
    const int k = idx / out_dim, o = idx % out_dim;
    dst[idx] = k < in_dim ? src[(size_t)o * in_dim + k] : 0.0f;
}

}  // namespace b200lap
