// Microbenchmarks behind the solver's step design (B200): dependent-chain latencies of shared/generic loads,
// store->load forwarding, barrier, shared atomics, FP64 / conversion pipe throughput, and the latency of fetching one
// random 8 KB matrix row with one 128-bit load per thread.  nvcc -arch=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_chain(int* out, long long* cyc, int iters, int* gbuf, int mode)
{
    extern __shared__ int sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) { sm[i] = (i * 97 + 13) & 4095; gbuf[i] = (i * 97 + 13) & 4095; }
    __syncthreads();
    int* gen = (mode & 1) ? (int*)sm : gbuf;      // generic pointer whose space the compiler cannot see
    if (mode & 2) gen = (int*)((size_t)gen ^ (size_t)out[0]);   // out[0] == 0: launder
    int idx = threadIdx.x & 4095;
    long long t0 = clock64();
    if (mode == 0) {           // LDS dependent chain
        for (int i = 0; i < iters; ++i) idx = sm[idx];
    } else if (mode == 3) {    // generic LD (to shared) dependent chain
        for (int i = 0; i < iters; ++i) idx = gen[idx];
    } else if (mode == 2) {    // generic LD (to global, L1) dependent chain
        for (int i = 0; i < iters; ++i) idx = gen[idx];
    } else if (mode == 4) {    // STS then dependent LDS of the same word
        for (int i = 0; i < iters; ++i) { sm[idx] = idx + 1; idx = sm[idx] & 4095; }
    } else if (mode == 5) {    // shared atomicAdd with the result used
        for (int i = 0; i < iters; ++i) idx = atomicAdd(&sm[idx & 31], 1) & 4095;
    } else if (mode == 6) {    // __syncthreads
        for (int i = 0; i < iters; ++i) __syncthreads();
    } else if (mode == 7) {    // generic ST then LD (shared)
        for (int i = 0; i < iters; ++i) { gen[idx] = idx + 1; idx = gen[idx] & 4095; }
    } else if (mode == 8) {    // dependent integer adds (ALU latency)
        for (int i = 0; i < iters; ++i) idx = idx * 3 + 1;
    } else if (mode == 9) {    // barrier + one LDS + branch (the minimal step skeleton)
        for (int i = 0; i < iters; ++i) { __syncthreads(); idx = sm[idx]; if (idx == 7777) break; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[1 + threadIdx.x] = idx;
}

__global__ void k_fp(double* out, long long* cyc, int iters, float* fin, int mode)
{
    float f0 = fin[threadIdx.x], f1 = fin[threadIdx.x + 1024], f2 = fin[threadIdx.x + 2048], f3 = fin[threadIdx.x + 3072];
    double a0 = f0, a1 = f1, a2 = f2, a3 = f3, s = 0.5;
    int cnt = 0;
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0) {          // 4 independent DADD chains
        for (int i = 0; i < iters; ++i) { a0 += s; a1 += s; a2 += s; a3 += s; }
    } else if (mode == 1) {   // F2F.F64.F32: 4 independent conversions per iteration
        for (int i = 0; i < iters; ++i) {
            a0 += (double)f0; a1 += (double)f1; a2 += (double)f2; a3 += (double)f3;
            f0 = __int_as_float(__float_as_int(f0) + 1); f1 = __int_as_float(__float_as_int(f1) + 1);
            f2 = __int_as_float(__float_as_int(f2) + 1); f3 = __int_as_float(__float_as_int(f3) + 1);
        }
    } else if (mode == 2) {   // DSETP: 4 compares per iteration feeding an integer counter
        for (int i = 0; i < iters; ++i) { cnt += (a0 < s) + (a1 < s) + (a2 < s) + (a3 < s); s = __longlong_as_double(__double_as_longlong(s) + 1); }
    } else if (mode == 3) {   // one dependent DADD chain (latency)
        for (int i = 0; i < iters; ++i) a0 += s;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = a0 + a1 + a2 + a3 + cnt;
}

// one CTA per "instance": fetch random rows of its own n x n float matrix, one float4 per thread, and reduce
__global__ void k_rows(const float* C, int n, int iters, long long* cyc, float* out, int stride_inst)
{
    const float* base = C + (size_t)blockIdx.x * stride_inst;
    __shared__ int s_next;
    unsigned rng = 12345u + blockIdx.x * 7919u;
    float acc = 0.f;
    long long tot = 0;
    int row = blockIdx.x % n;
    for (int i = 0; i < iters; ++i) {
        long long t0 = clock64();
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * n) + threadIdx.x);
        acc += v.x + v.y + v.z + v.w;
        if (acc == 12345.678f) out[0] = acc;          // force the wait here
        long long t1 = clock64();
        tot += t1 - t0;
        rng = rng * 1664525u + 1013904223u;
        if (threadIdx.x == 0) s_next = (rng >> 8) % n;
        __syncthreads();
        row = s_next;
        __syncthreads();
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = tot / iters;
    out[1 + blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
    int *out, *gbuf; long long* cyc; double* dout; float* fin;
    CK(cudaMalloc(&out, 8192 * 4)); CK(cudaMemset(out, 0, 8192 * 4)); CK(cudaMalloc(&gbuf, 4096 * 4)); CK(cudaMalloc(&cyc, 1024 * 8));
    CK(cudaMalloc(&dout, 1024 * 8)); CK(cudaMalloc(&fin, 4096 * 4)); CK(cudaMemset(fin, 0x3f, 4096 * 4));
    long long h[1024];
    const char* names[] = {"LDS chain", "", "generic LD -> global(L1) chain", "generic LD -> shared chain", "STS+LDS same word", "ATOMS.ADD (result used)",
                           "__syncthreads", "generic ST+LD same word (shared)", "IMAD chain", "barrier + LDS + branch"};
    const int iters = 2000;
    for (int T : {32, 512}) {
        for (int mode : {0, 2, 3, 4, 5, 6, 7, 8, 9}) {
            k_chain<<<1, T, 16384>>>(out, cyc, iters, gbuf, mode);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("T=%4d %-36s %7.1f cycles/iter\n", T, names[mode], (double)h[0] / iters);
        }
    }
    const char* fnames[] = {"DADD x4 independent", "F2F.F64.F32 x4 (+DADD x4)", "DSETP x4", "DADD dependent chain"};
    for (int T : {32, 128, 512}) {
        for (int mode : {0, 1, 2, 3}) {
            k_fp<<<1, T>>>(dout, cyc, iters, fin, mode);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("T=%4d %-36s %7.1f cycles/iter (%d warps on one SM)\n", T, fnames[mode], (double)h[0] / iters, T / 32);
        }
    }
    // random-row fetch latency: B instances of 2048 x 2048 floats (16 MB each)
    const int n = 2048;
    for (int B : {1, 7, 64}) {
        float* C; float* fo;
        CK(cudaMalloc(&C, (size_t)B * n * n * 4)); CK(cudaMemset(C, 0, (size_t)B * n * n * 4)); CK(cudaMalloc(&fo, (size_t)(1 + B * 512) * 4));
        k_rows<<<B, 512>>>(C, n, 3000, cyc, fo, n * n);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, 8 * B, cudaMemcpyDeviceToHost));
        double s = 0; for (int b = 0; b < B; ++b) s += h[b];
        printf("random 8 KB row fetch, %2d concurrent instances of 16 MB: %7.0f cycles per row (issue -> all 4 words usable, thread 0)\n", B, s / B);
        CK(cudaFree(C)); CK(cudaFree(fo));
    }
    return 0;
}
