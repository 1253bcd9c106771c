// api.cu -- host orchestration + the C ABI of libb200lap.so (include/b200lap.h).
//
// Every entry point enqueues hand-written sm_100a kernels on the context's stream; nothing here
// computes on the host.  Built by nvcc only (build.py); the same translation unit also compiles
// against tests/emul/cuda_emul.h for the CPU-side logic tests (test infrastructure, not shipped).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "colsweep.cuh"
#include "frontend.cuh"
#include "solver.cuh"
#include "dualsweep.cuh"
#include "features.cuh"
#include "features_smem.cuh"
#include "features_warp.cuh"
#include "features_group.cuh"
#include "mlp.cuh"
#include "mlp_tc.cuh"
#include "../../include/b200lap.h"

// host_narrow.cpp: binary64 -> binary32 marshalling of host buffers (multi-threaded, exactness-checked)
namespace b200lap_host {
int narrow_threads();
int narrow_percent();
bool narrow(const double* src, float* dst, size_t count, int threads);
}

using namespace b200lap;

// ---------------------------------------------------------------------------------------------
// context, errors, workspace
// ---------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(B200LAP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

}  // namespace

struct b200lap_ws_block {
    unsigned char* p;
    size_t size, off;
};
typedef b200lap_ws_block WsBlock;

struct b200lap_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::vector<WsBlock> blocks;
    long long launches = 0;
    int max_dyn_smem = 227 * 1024 - 4096;
    int solver_threads = 0;      // option: force the solver block size (tests)
    int force_global_state = 0;  // option: keep solver state in global memory (tests)
    int solver_cluster = 0;      // option: CTAs per instance (thread-block cluster), 0 = auto, 1 = single CTA
    int solver_cluster_min_n = 8192;   // option: auto mode uses a cluster from this size on
    int solver_regpath = 1;      // option: augmentation with register-resident d/v (solver_path.cuh); 0 = shared-memory state path
    int solver_kcap = 0;         // option: scans per batched relax step of the register path (0 = auto)
    int solver_pipe = 1;         // option: replay a batch's hits behind the next batch's row fetch
    int front_rows_per_cta = 0;  // option
    int mlp_impl = 0;            // option: 0 = default
    int feat_ept = 0;            // option: entries per thread of the row-feature kernel (0 = auto)
    int feat_impl = 0;           // option: 0 = shared-memory streaming kernel for binary32 storage, 1 = register-resident kernel
    int feat_threads = 0;        // options of the streaming kernel (0 = auto): CTA size, row buffers, sample size, CTAs per SM
    int feat_nbuf = 0;
    int feat_nsamp = 0;
    int feat_ctas = 0;
    const int* feat_redo_count = nullptr;   // device counter of the last row-feature call (rows the fast kernel handed to the fall-back)
    cudaStream_t feat_redo_stream = nullptr;
    int feat_stream = 0;         // option: 0 = auto, 1 = stream the row from global memory (no staging), 2 = stage it in shared memory
    int feat_torch_mode = 0;     // option: 1 = the definitions of compute_row_features_torch (gnn/features.py:246-351)
    int feat_group = 0;          // option: warps per row of the group kernel (0 = auto: n / 2048, 1..8 forces; needs n = 32 * G * {16, 32, 64, 128})
    int sm_count = 148;
    std::map<int, float*> posenc;   // positional-encoding tables [n][8], one per n ever used (never freed before the ctx)
    std::mutex mu;
    // Second LANE (its own stream and workspace blocks): independent whole-pipeline calls alternate between the two
    // lanes when `overlap_steps` is on, so two batches are in flight and the SMs a 64-instance batch leaves idle
    // (one CTA per instance) work on the other batch.  `stream` / `blocks` always name the ACTIVE lane; outside a
    // call that is lane 0.
    static constexpr int kMaxLanes = 8;
    struct LaneStore { cudaStream_t stream = nullptr; std::vector<WsBlock> blocks; cudaEvent_t ev = nullptr; };
    LaneStore parked[kMaxLanes];  // lanes 1..: their stream/blocks while inactive; slot a holds LANE 0's while lane a is active
    int active_lane = 0;
    int overlap_steps = 0;       // option: number of lanes independent whole-pipeline calls rotate through (0/1 = off, up to 8)
    long long lane_calls = 0;
    int last_lane = 0;           // lane of the most recent whole-pipeline call (b200lap_ctx_last_lane)
    int lanes() const { return overlap_steps < 2 ? 1 : (overlap_steps > kMaxLanes ? kMaxLanes : overlap_steps); }
    void use_lane(int l) {
        if (l == active_lane) return;
        if (active_lane != 0) { std::swap(stream, parked[active_lane].stream); std::swap(blocks, parked[active_lane].blocks); active_lane = 0; }
        if (l != 0) { std::swap(stream, parked[l].stream); std::swap(blocks, parked[l].blocks); active_lane = l; }
    }
    cudaStream_t lane_stream(int l) const {
        if (l == active_lane) return stream;
        return l == 0 ? parked[active_lane].stream : parked[l].stream;
    }

    void ws_reset() {
        for (auto& b : blocks) b.off = 0;
    }
    void* take(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        for (auto& b : blocks)
            if (b.size - b.off >= bytes) {
                void* r = b.p + b.off;
                b.off += bytes;
                return r;
            }
        size_t sz = bytes > ((size_t)32 << 20) ? bytes : ((size_t)32 << 20);
        void* p = nullptr;
        if (cudaMalloc(&p, sz) != cudaSuccess) {
            (void)cudaGetLastError();      // a failed cudaMalloc leaves the error set: do not let a later launch check see it
            // a smaller exact-size block may still fit
            sz = bytes;
            if (cudaMalloc(&p, sz) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        }
        blocks.push_back(WsBlock{(unsigned char*)p, sz, bytes});
        return p;
    }
    template <typename T> T* take_n(size_t count) { return (T*)take(count * sizeof(T)); }
};

#define TAKE(var, type, count)                                                         \
    type* var = ctx->take_n<type>(count);                                              \
    if (!var) return fail(-1, "device workspace allocation failed (" #var ")")

namespace {

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

template <typename CT> constexpr int natural_vec() { return sizeof(CT) == 4 ? 4 : 2; }

template <typename CT> bool vec_ok(const CT* C, long long inst_stride, int ld, int n) {
    const int V = natural_vec<CT>();
    return n % V == 0 && ld % V == 0 && inst_stride % V == 0 && ((uintptr_t)C % 16) == 0;
}

// ---- column sweeps ------------------------------------------------------------------------------
inline int strip_rows(int n) {
    int r = n / 64;
    if (r < 32) r = 32;
    if (r > 256) r = 256;
    return r;
}

template <typename CT>
int run_col_argmin(b200lap_ctx* ctx, const CT* C, long long inst_stride, int ld, int batch, int n, CT* colmin, int* colarg)
{
    const int rps = strip_rows(n);
    const int S = (n + rps - 1) / rps;
    TAKE(pval, CT, (size_t)batch * S * n);
    TAKE(prow, int, (size_t)batch * S * n);
    if (vec_ok(C, inst_stride, ld, n)) {
        constexpr int V = natural_vec<CT>();
        dim3 grid((n + kColThreads * V - 1) / (kColThreads * V), S, batch);
        auto k = k_col_argmin_partial<CT, V>;
        B200LAP_LAUNCH(k, grid, dim3(kColThreads), 0, ctx->stream, C, inst_stride, ld, n, rps, pval, prow);
    } else {
        dim3 grid((n + kColThreads - 1) / kColThreads, S, batch);
        auto k = k_col_argmin_partial<CT, 1>;
        B200LAP_LAUNCH(k, grid, dim3(kColThreads), 0, ctx->stream, C, inst_stride, ld, n, rps, pval, prow);
    }
    {
        dim3 grid((n + 255) / 256, batch);
        auto k = k_col_argmin_final<CT>;
        B200LAP_LAUNCH(k, grid, dim3(256), 0, ctx->stream, (const CT*)pval, (const int*)prow, S, n, colmin, colarg);
    }
    ctx->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

template <typename CT>
int run_min_trick(b200lap_ctx* ctx, const CT* C, long long inst_stride, int ld, int batch, int n, const float* u, double* v)
{
    const int rps = strip_rows(n);
    const int S = (n + rps - 1) / rps;
    TAKE(pval, double, (size_t)batch * S * n);
    if (vec_ok(C, inst_stride, ld, n)) {
        constexpr int V = natural_vec<CT>();
        dim3 grid((n + kColThreads * V - 1) / (kColThreads * V), S, batch);
        auto k = k_min_trick_partial<CT, V>;
        B200LAP_LAUNCH(k, grid, dim3(kColThreads), 0, ctx->stream, C, inst_stride, ld, n, rps, u, pval);
    } else {
        dim3 grid((n + kColThreads - 1) / kColThreads, S, batch);
        auto k = k_min_trick_partial<CT, 1>;
        B200LAP_LAUNCH(k, grid, dim3(kColThreads), 0, ctx->stream, C, inst_stride, ld, n, rps, u, pval);
    }
    {
        dim3 grid((n + 255) / 256, batch);
        B200LAP_LAUNCH(k_min_trick_final, grid, dim3(256), 0, ctx->stream, (const double*)pval, S, n, v);
    }
    ctx->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

// ---- front end ------------------------------------------------------------------------------------
__global__ void k_force_slow_path(FrontFlags* flags, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) flags[b].any_viol = 1;
}

template <typename CT, int VEC, int EPT, bool VSM, int MAXT>
int launch_front(b200lap_ctx* ctx, int rows_per_cta, const CT* C, long long inst_stride, int ld, int batch, int n,
                 const double* u, const double* v, double eps, double tol, double* u_tight, int* tl, int* tc, FrontFlags* flags)
{
    const int T = round_up((n + EPT - 1) / EPT, 32);
    dim3 grid((n + rows_per_cta - 1) / rows_per_cta, batch);
    auto k = k_front_end<CT, VEC, EPT, VSM, MAXT>;
    const size_t smem = VSM ? (size_t)n * sizeof(double) : 0;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200LAP_LAUNCH(k, grid, dim3(T), smem, ctx->stream, C, inst_stride, ld, n, rows_per_cta, u, v, eps, tol, u_tight, tl, tc, flags);
    ctx->launches += 1;
    return 0;
}

// Few warps per row with many entries per thread: the per-row block reduction and loop overhead are
// per-thread costs (ncu: the 4-entries-per-thread version issued 62 instructions per matrix entry).
template <typename CT, int VEC>
int dispatch_front(b200lap_ctx* ctx, int R, const CT* C, long long inst_stride, int ld, int batch, int n, const double* u,
                   const double* v, double eps, double tol, double* u_tight, int* tl, int* tc, FrontFlags* flags, bool* done)
{
#define FRONT(EPT_, VSM_, MAXT_) \
    return launch_front<CT, VEC, EPT_, VSM_, MAXT_>(ctx, R, C, inst_stride, ld, batch, n, u, v, eps, tol, u_tight, tl, tc, flags)
    *done = true;
    if (n <= 1024) FRONT(4, false, 256);
    if (n <= 4096) FRONT(16, false, 256);
    if (n <= 8192) FRONT(32, true, 256);
    if (n <= 16384) FRONT(32, true, 512);
#undef FRONT
    *done = false;
    return 0;
}

template <typename CT>
int run_front_end(b200lap_ctx* ctx, const CT* C, long long inst_stride, int ld, int batch, int n, const double* u,
                  const double* v, double eps, double* u_tight, int* tl, int* tc, FrontFlags* flags)
{
    const double tol = eps > 1e-9 ? eps : 1e-9;
    CK(cudaMemsetAsync(flags, 0, sizeof(FrontFlags) * (size_t)batch, ctx->stream));
    int R = ctx->front_rows_per_cta > 0 ? ctx->front_rows_per_cta : (n >= 8192 ? 16 : (n >= 1024 ? 8 : 2));
    const bool vec = vec_ok(C, inst_stride, ld, n);
    constexpr int V = natural_vec<CT>();
    bool done = false;
    int fr = vec ? dispatch_front<CT, V>(ctx, R, C, inst_stride, ld, batch, n, u, v, eps, tol, u_tight, tl, tc, flags, &done)
                 : dispatch_front<CT, 1>(ctx, R, C, inst_stride, ld, batch, n, u, v, eps, tol, u_tight, tl, tc, flags, &done);
    if (fr) return fr;
    if (!done) {
        // row too long for the register-resident sweep: the solver kernel runs the (sequential)
        // front end itself -- still on the device
        B200LAP_LAUNCH(k_force_slow_path, dim3((batch + 127) / 128), dim3(128), 0, ctx->stream, flags, batch);
        ctx->launches += 1;
    }
    CK(cudaGetLastError());
    return 0;
}

// ---- solve ----------------------------------------------------------------------------------------
// launch `batch` thread-block clusters of `cluster` CTAs each (one cluster per instance)
template <typename K, typename A>
cudaError_t launch_clustered(K kernel, int batch, int cluster, int T, size_t smem, cudaStream_t stream, const A& args)
{
#ifdef B200LAP_EMUL
    (void)kernel; (void)batch; (void)cluster; (void)T; (void)smem; (void)stream; (void)args;
    return cudaErrorInvalidValue;
#else
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)batch * (unsigned)cluster);
    cfg.blockDim = dim3((unsigned)T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args);
#endif
}

template <typename CT>
int run_solve(b200lap_ctx* ctx, const CT* C, long long inst_stride, int ld, int batch, int n, const double* u_seed,
              const double* v_seed, double eps, int mode, int* x, int* y, int* rc, long long* trace, double* v_out,
              const CT* colmin_in = nullptr, const int* colarg_in = nullptr)
{
    if (batch <= 0) return 0;
    if (n <= 0) return fail(-2, "n <= 0");
    // column minima + first rows: handed over by the dense pass when it ran on the same matrices (no fifth sweep of C)
    const CT* colmin = colmin_in;
    const int* colarg = colarg_in;
    int r = 0;
    if (!colmin || !colarg) {
        TAKE(cm, CT, (size_t)batch * n);
        TAKE(ca, int, (size_t)batch * n);
        r = run_col_argmin(ctx, C, inst_stride, ld, batch, n, cm, ca);
        if (r) return r;
        colmin = cm; colarg = ca;
    }
    SolveArgs<CT> a;
    a.C = C; a.inst_stride = inst_stride; a.ld = ld; a.n = n;
    a.u_seed = u_seed; a.v_seed = v_seed; a.eps = eps; a.mode = mode;
    a.u_tight = nullptr; a.tight_cols = nullptr; a.tight_cnt = nullptr; a.flags = nullptr;
    if (mode == 0) {
        TAKE(u_tight, double, (size_t)batch * n);
        TAKE(tl, int, (size_t)batch * n * kTightCap);
        TAKE(tc, int, (size_t)batch * n);
        TAKE(flags, FrontFlags, (size_t)batch);
        r = run_front_end(ctx, C, inst_stride, ld, batch, n, u_seed, v_seed, eps, u_tight, tl, tc, flags);
        if (r) return r;
        a.u_tight = u_tight; a.tight_cols = tl; a.tight_cnt = tc; a.flags = flags;
    }
    a.colmin = colmin; a.colarg = colarg;
    const size_t state = solver_state_bytes(n);
    size_t smem = 0;
    // cluster mode (solver.cuh): relax steps spread over the CTAs of a thread-block cluster, state in the global workspace
    int cluster = ctx->solver_cluster;
    if (cluster < 0) cluster = 0;
    // auto: 8 CTAs measured best at n = 8192 and 16384 (tools/sweep_cluster.py); a cold solve is all column reduction
    // and ARR on the master, where the cluster only costs (5.6 -> 6.5 s per 8 x 8192 batch), so it stays on one CTA
    if (cluster == 0) cluster = (n >= ctx->solver_cluster_min_n && mode == 0) ? 8 : 1;
    if (cluster > 8) cluster = 8;
    if (cluster > 1 && n < 32 * cluster) cluster = 1;
#ifdef B200LAP_EMUL
    cluster = 1;
#endif
    a.smem_mask = solver_place_state(n, ctx->force_global_state ? 0 : (size_t)ctx->max_dyn_smem, &smem,
                                     cluster > 1 ? kClusterSmemArrays : (1 << ST_COUNT) - 1);
    a.gws = nullptr; a.gws_stride = 0;
    if (a.smem_mask != (1 << ST_COUNT) - 1) {
        // (+ the per-scan column bitmaps of the cluster mode's batched relax step, at the tail of the stride)
        const size_t cbm_bytes = (size_t)kMaxScans * ((((size_t)n + 31) / 32 + 4 + 3) & ~(size_t)3) * 4;
        const size_t stride = (state + 4096 + cbm_bytes + 255) & ~(size_t)255;
        unsigned char* g = (unsigned char*)ctx->take(stride * (size_t)batch);
        if (!g) return fail(-1, "device workspace allocation failed (solver state)");
        a.gws = g; a.gws_stride = (long long)stride;
    }
    a.x = x; a.y = y; a.rc = rc; a.trace = trace; a.v_out = v_out;
    a.cluster = cluster;
    int T = ctx->solver_threads > 0 ? ctx->solver_threads : round_up((n + 3) / 4, 32);
    if (T > 1024) T = 1024;
    if (T < 32) T = 32;
    // cluster mode: 512-thread CTAs synchronise faster and the 128-register variants do not spill; measured 1.46 s
    // against 1.66 s for 4 x 8192 (tools/sweep_cluster_threads.py); above 8192 the master needs all 1024 threads
    if (cluster > 1 && ctx->solver_threads <= 0 && n <= 8192 && T > 512) T = 512;
    // register-resident augmentation (solver_path.cuh): 128-bit row loads, d/v of a thread's columns in registers.
    // 512-thread CTAs keep the per-thread state (4 registers per column) under the 128-register budget.
    a.regpath = (ctx->solver_regpath && cluster == 1 && vec_ok(C, inst_stride, ld, n)) ? 1 : 0;
    if (a.regpath && ctx->solver_threads <= 0 && T > 512 && n <= 8192) T = 512;
    // (256-thread CTAs -- two instances per SM -- were measured for the many-batches-in-flight case: 27.9 vs 27.8 ms per
    //  64 x 2048 step at 8 lanes, but 113 vs 83 ms for a batch alone and a slower host path; not used.  tools/solver_threads_exp.py)
    a.kcap = ctx->solver_kcap; a.pipe = ctx->solver_pipe;
    const int per_thread = (n + T - 1) / T;      // row entries a thread keeps in registers per step
#define SOLVE(MAXC_)                                                                                                       \
    do {                                                                                                                   \
        auto k = T <= 512 ? k_solve<CT, MAXC_, 512> : k_solve<CT, MAXC_, 1024>;                                            \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024))); \
        bool clustered = cluster > 1;                                                                                      \
        if (clustered && launch_clustered(k, batch, cluster, T, smem, ctx->stream, a) != cudaSuccess) {                    \
            /* a cluster shape the device refuses: same kernel, one CTA per instance (the state placement stays valid) */  \
            (void)cudaGetLastError();                                                                                      \
            a.cluster = 1; clustered = false;                                                                              \
        }                                                                                                                  \
        if (!clustered) B200LAP_LAUNCH(k, dim3(batch), dim3(T), smem, ctx->stream, a);                                     \
    } while (0)
    if (per_thread <= 4) SOLVE(4);
    else if (per_thread <= 8) SOLVE(8);
    else if (per_thread <= 16) SOLVE(16);
    else SOLVE(0);
#undef SOLVE
    ctx->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

__global__ void k_narrow(const double* __restrict__ src, long long count, float* __restrict__ dst, int* exact) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    int bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const double c = src[i];
        const float f = (float)c;
        dst[i] = f;
        bad |= !((double)f == c);
    }
    if (bad) atomicAnd(exact, 0);
}

// int32 assignments -> int64 (the reference's long long x, y); rows of instances with rc != 0 read -1
__global__ void k_widen_ids(const int* __restrict__ src, const int* __restrict__ rc, int n, long long count, long long* __restrict__ dst) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        dst[i] = rc[i / n] == 0 ? (long long)src[i] : -1ll;
}

int run_narrow(b200lap_ctx* ctx, const double* src, long long count, float* dst, int* exact) {
    long long blocks = (count + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    B200LAP_LAUNCH(k_narrow, dim3((unsigned)blocks), dim3(256), 0, ctx->stream, src, count, dst, exact);
    ctx->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

std::mutex g_default_mu;
b200lap_ctx* g_default = nullptr;

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* b200lap_last_error(void) { return g_err.c_str(); }

int b200lap_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
    return c;
}

int b200lap_ctx_create(int device, void* stream, b200lap_ctx** out) {
    if (!out) return fail(B200LAP_ERR_ARG, "out is null");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return fail(B200LAP_ERR_CUDA, "no CUDA device is visible: libb200lap has no CPU path");
    if (device < 0 || device >= count) return fail(B200LAP_ERR_ARG, "device index out of range");
    CK(cudaSetDevice(device));
    b200lap_ctx* c = new (std::nothrow) b200lap_ctx();
    if (!c) return fail(-1, "host allocation failed");
    c->device = device;
#ifndef B200LAP_EMUL
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) c->sm_count = sms;
    }
#endif
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail(B200LAP_ERR_CUDA, "cudaStreamCreate failed"); }
        c->own_stream = true;
    }
    for (int l = 1; l < b200lap_ctx::kMaxLanes; ++l) {
        if (cudaStreamCreateWithFlags(&c->parked[l].stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->parked[l].ev, cudaEventDisableTiming) != cudaSuccess) {
            (void)cudaGetLastError();
            delete c;
            return fail(B200LAP_ERR_CUDA, "cudaStreamCreate failed (lanes)");
        }
    }
    *out = c;
    return 0;
}

void b200lap_ctx_destroy(b200lap_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->use_lane(0);
    cudaStreamSynchronize(ctx->stream);
    for (auto& b : ctx->blocks) cudaFree(b.p);
    for (int l = 1; l < b200lap_ctx::kMaxLanes; ++l) {
        if (ctx->parked[l].stream) cudaStreamSynchronize(ctx->parked[l].stream);
        for (auto& b : ctx->parked[l].blocks) cudaFree(b.p);
        if (ctx->parked[l].ev) cudaEventDestroy(ctx->parked[l].ev);
        if (ctx->parked[l].stream) cudaStreamDestroy(ctx->parked[l].stream);
    }
    for (auto& kv : ctx->posenc) cudaFree(kv.second);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void* b200lap_ctx_stream(b200lap_ctx* ctx) { return ctx ? (void*)ctx->lane_stream(0) : nullptr; }
int b200lap_ctx_last_lane(b200lap_ctx* ctx) { return ctx ? ctx->last_lane : 0; }

void* b200lap_ctx_lane_stream(b200lap_ctx* ctx, int lane) {
    return (ctx && lane >= 0 && lane < b200lap_ctx::kMaxLanes) ? (void*)ctx->lane_stream(lane) : nullptr;
}

int b200lap_ctx_sync(b200lap_ctx* ctx) {
    if (!ctx) return fail(B200LAP_ERR_ARG, "ctx is null");
    for (int l = 0; l < b200lap_ctx::kMaxLanes; ++l) CK(cudaStreamSynchronize(ctx->lane_stream(l)));
    return 0;
}

/* lane 0's stream waits (on the device) for everything enqueued on the other lanes so far */
int b200lap_ctx_join(b200lap_ctx* ctx) {
    if (!ctx) return fail(B200LAP_ERR_ARG, "ctx is null");
    for (int l = 1; l < b200lap_ctx::kMaxLanes; ++l) {
        CK(cudaEventRecord(ctx->parked[l].ev, ctx->lane_stream(l)));
        CK(cudaStreamWaitEvent(ctx->lane_stream(0), ctx->parked[l].ev, 0));
    }
    return 0;
}

long long b200lap_ctx_launch_count(b200lap_ctx* ctx) { return ctx ? ctx->launches : 0; }

long long b200lap_ctx_feature_redo_rows(b200lap_ctx* ctx) {
    if (!ctx || !ctx->feat_redo_count) return 0;
    int c[4] = {0, 0, 0, 0};
    if (cudaStreamSynchronize(ctx->feat_redo_stream) != cudaSuccess) return -1;
    if (cudaMemcpy(c, ctx->feat_redo_count, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (getenv("B200LAP_FEAT_DEBUG")) fprintf(stderr, "[b200lap] feature redo rows %d: list overflow / candidates %d, bracket miss %d, tie-heavy target bin %d\n", c[0], c[1], c[2], c[3]);
    return c[0];
}

int b200lap_ctx_set_option(b200lap_ctx* ctx, const char* key, long long value) {
    if (!ctx || !key) return fail(B200LAP_ERR_ARG, "null argument");
    const std::string k(key);
    if (k == "solver_threads") ctx->solver_threads = (int)value;
    else if (k == "force_global_state") ctx->force_global_state = (int)value;
    else if (k == "solver_cluster") ctx->solver_cluster = (int)value;
    else if (k == "solver_cluster_min_n") ctx->solver_cluster_min_n = (int)value;
    else if (k == "solver_regpath") ctx->solver_regpath = (int)value;
    else if (k == "solver_kcap") ctx->solver_kcap = (int)value;
    else if (k == "solver_pipe") ctx->solver_pipe = (int)value;
    else if (k == "overlap_steps") ctx->overlap_steps = (int)value;
    else if (k == "solver_smem_budget") ctx->max_dyn_smem = value > 0 ? (int)value : 227 * 1024 - 4096;
    else if (k == "front_rows_per_cta") ctx->front_rows_per_cta = (int)value;
    else if (k == "mlp_impl") ctx->mlp_impl = (int)value;
    else if (k == "feat_ept") ctx->feat_ept = (int)value;
    else if (k == "feat_impl") ctx->feat_impl = (int)value;
    else if (k == "feat_threads") ctx->feat_threads = (int)value;
    else if (k == "feat_nbuf") ctx->feat_nbuf = (int)value;
    else if (k == "feat_nsamp") ctx->feat_nsamp = (int)value;
    else if (k == "feat_ctas") ctx->feat_ctas = (int)value;
    else if (k == "feat_group") ctx->feat_group = (int)value;
    else if (k == "feat_stream") ctx->feat_stream = (int)value;
    else if (k == "feat_torch_mode") ctx->feat_torch_mode = (int)value;
    else return fail(B200LAP_ERR_ARG, "unknown option " + k);
    return 0;
}

b200lap_ctx* b200lap_default_ctx(void) {
    std::lock_guard<std::mutex> g(g_default_mu);
    if (!g_default) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        b200lap_ctx* c = nullptr;
        if (b200lap_ctx_create(dev, nullptr, &c) != 0) return nullptr;
        g_default = c;
    }
    return g_default;
}

int b200lap_dev_narrow(b200lap_ctx* ctx, const double* src, long long count, float* dst, int* exact) {
    if (!ctx) return fail(B200LAP_ERR_ARG, "ctx is null");
    return run_narrow(ctx, src, count, dst, exact);
}

int b200lap_dev_col_argmin(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, void* colmin, int* colarg) {
    if (!ctx) return fail(B200LAP_ERR_ARG, "ctx is null");
    if (n <= 0 || batch <= 0) return fail(B200LAP_ERR_ARG, "empty problem");
    ctx->ws_reset();
    const long long st = (long long)n * n;
    return is_f64 ? run_col_argmin(ctx, (const double*)C, st, n, batch, n, (double*)colmin, colarg)
                  : run_col_argmin(ctx, (const float*)C, st, n, batch, n, (float*)colmin, colarg);
}

int b200lap_dev_min_trick(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, const float* u, double* v) {
    if (!ctx) return fail(B200LAP_ERR_ARG, "ctx is null");
    if (n <= 0 || batch <= 0) return fail(B200LAP_ERR_ARG, "empty problem");
    ctx->ws_reset();
    const long long st = (long long)n * n;
    return is_f64 ? run_min_trick(ctx, (const double*)C, st, n, batch, n, u, v)
                  : run_min_trick(ctx, (const float*)C, st, n, batch, n, u, v);
}

int b200lap_dev_solve_seeded(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, const double* u_seed,
                             const double* v_seed, double eps, int* x, int* y, int* rc, long long* trace, double* v_out) {
    if (!ctx) return fail(B200LAP_ERR_ARG, "ctx is null");
    ctx->ws_reset();
    const long long st = (long long)n * n;
    return is_f64 ? run_solve(ctx, (const double*)C, st, n, batch, n, u_seed, v_seed, eps, 0, x, y, rc, trace, v_out)
                  : run_solve(ctx, (const float*)C, st, n, batch, n, u_seed, v_seed, eps, 0, x, y, rc, trace, v_out);
}

int b200lap_dev_solve_cold(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, int* x, int* y, int* rc,
                           long long* trace, double* v_out) {
    if (!ctx) return fail(B200LAP_ERR_ARG, "ctx is null");
    ctx->ws_reset();
    const long long st = (long long)n * n;
    return is_f64 ? run_solve(ctx, (const double*)C, st, n, batch, n, nullptr, nullptr, 0.0, 1, x, y, rc, trace, v_out)
                  : run_solve(ctx, (const float*)C, st, n, batch, n, nullptr, nullptr, 0.0, 1, x, y, rc, trace, v_out);
}

int b200lap_dev_front_end(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, const double* u_seed,
                          const double* v_seed, double eps, double* u_tight, int* tight_cnt, int* flags) {
    if (!ctx) return fail(B200LAP_ERR_ARG, "ctx is null");
    if (n <= 0 || batch <= 0) return fail(B200LAP_ERR_ARG, "empty problem");
    static_assert(sizeof(FrontFlags) == 16, "flags are exported as 4 x int32");
    ctx->ws_reset();
    TAKE(tl, int, (size_t)batch * n * kTightCap);
    const long long st = (long long)n * n;
    return is_f64 ? run_front_end(ctx, (const double*)C, st, n, batch, n, u_seed, v_seed, eps, u_tight, tl, tight_cnt, (FrontFlags*)flags)
                  : run_front_end(ctx, (const float*)C, st, n, batch, n, u_seed, v_seed, eps, u_tight, tl, tight_cnt, (FrontFlags*)flags);
}

}  // extern "C"

// ---- dual-potential sweeps (solvers/advanced_dual.py:14-63) ---------------------------------------------
namespace {

template <typename CT>
int run_col_min_reduced(b200lap_ctx* ctx, const CT* C, long long inst_stride, int ld, int batch, int n, const double* u, double* v_cap)
{
    const int rps = strip_rows(n);
    const int S = (n + rps - 1) / rps;
    TAKE(pval, double, (size_t)batch * S * n);
    if (vec_ok(C, inst_stride, ld, n)) {
        constexpr int V = natural_vec<CT>();
        dim3 grid((n + kColThreads * V - 1) / (kColThreads * V), S, batch);
        auto k = k_col_min_reduced_partial<CT, V>;
        B200LAP_LAUNCH(k, grid, dim3(kColThreads), 0, ctx->stream, C, inst_stride, ld, n, rps, u, pval);
    } else {
        dim3 grid((n + kColThreads - 1) / kColThreads, S, batch);
        auto k = k_col_min_reduced_partial<CT, 1>;
        B200LAP_LAUNCH(k, grid, dim3(kColThreads), 0, ctx->stream, C, inst_stride, ld, n, rps, u, pval);
    }
    B200LAP_LAUNCH(k_min_trick_final, dim3((n + 255) / 256, batch), dim3(256), 0, ctx->stream, (const double*)pval, S, n, v_cap);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

// min_ij ((c - u_i) - v_j) per instance -> host doubles; optionally writes the reduced-cost matrices
template <typename CT>
int run_reduced_costs(b200lap_ctx* ctx, const CT* C, long long inst_stride, int ld, int batch, int n, const double* u, const double* v,
                      double* out, double* min_host)
{
    TAKE(mo, unsigned long long, (size_t)batch);
    CK(cudaMemsetAsync(mo, 0xff, sizeof(unsigned long long) * (size_t)batch, ctx->stream));
    int gx = n < ctx->sm_count * 4 ? n : ctx->sm_count * 4;
    if (out) { auto k = k_reduced_costs<CT, true>; B200LAP_LAUNCH(k, dim3(gx, batch), dim3(256), 0, ctx->stream, C, inst_stride, ld, n, u, v, out, mo); }
    else { auto k = k_reduced_costs<CT, false>; B200LAP_LAUNCH(k, dim3(gx, batch), dim3(256), 0, ctx->stream, C, inst_stride, ld, n, u, v, out, mo); }
    ctx->launches += 1;
    CK(cudaGetLastError());
    std::vector<unsigned long long> h((size_t)batch);
    CK(cudaMemcpyAsync(h.data(), mo, sizeof(unsigned long long) * (size_t)batch, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < batch; ++b) {
        const unsigned long long k = h[(size_t)b];
        const long long bits = (long long)(k ^ (((long long)k < 0) ? 0x8000000000000000ull : 0xffffffffffffffffull));
        double d;
        static_assert(sizeof(d) == sizeof(bits), "binary64");
        memcpy(&d, &bits, sizeof(d));
        min_host[b] = k == ~0ull ? INFINITY : d;
    }
    return 0;
}

// project_feasible on device buffers of ONE instance (u, v updated in place) -> rounds used
template <typename CT>
int run_project_feasible(b200lap_ctx* ctx, const CT* C, int n, double* u, double* v, int max_rounds, double tol, int* rounds_out)
{
    const long long st = (long long)n * n;
    TAKE(cap, double, (size_t)n);
    TAKE(tl, int, (size_t)n * kTightCap);
    TAKE(tc, int, (size_t)n);
    TAKE(flags, FrontFlags, 1);
    const int rounds = max_rounds < 1 ? 1 : max_rounds;
    int used = 0;
    // u_cap_i = min_j (c_ij - v_j) is the front-end sweep's row tightening; the same sweep evaluates the reference's
    // stopping rule any((c - u_i) - v_j < -tol) for the potentials it is handed
    int r = run_front_end(ctx, C, st, n, 1, n, u, v, tol, cap, tl, tc, flags);
    if (r) return r;
    for (int round = 0; round < rounds; ++round) {
        ++used;
        B200LAP_LAUNCH(k_clamp_min, dim3((n + 255) / 256), dim3(256), 0, ctx->stream, u, (const double*)cap, (long long)n);
        r = run_col_min_reduced(ctx, C, st, n, 1, n, u, cap);
        if (r) return r;
        B200LAP_LAUNCH(k_clamp_min, dim3((n + 255) / 256), dim3(256), 0, ctx->stream, v, (const double*)cap, (long long)n);
        ctx->launches += 2;
        r = run_front_end(ctx, C, st, n, 1, n, u, v, tol, cap, tl, tc, flags);     // feasibility of (u, v) + next round's u_cap
        if (r) return r;
        FrontFlags hf;
        CK(cudaMemcpyAsync(&hf, flags, sizeof(FrontFlags), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (!hf.infeasible) break;
    }
    if (rounds_out) *rounds_out = used;
    return 0;
}


// Oracle duals by difference constraints (solvers/dual_computation.py:13-47): Jacobi rounds of the Bellman-Ford
// relaxation from v = 0 until no column potential moves (checked on the host every few rounds) or n - 1 rounds.
template <typename CT>
int run_bf_duals(b200lap_ctx* ctx, const CT* C, int batch, int n, const int* x, double* v, int* rounds_out)
{
    const long long st = (long long)n * n;
    const int rps = strip_rows(n);
    const int S = (n + rps - 1) / rps;
    const size_t cnt = (size_t)batch * n;
    TAKE(pval, double, (size_t)batch * S * n);
    TAKE(cand, double, cnt);
    TAKE(vx, double, cnt);
    TAKE(cx, double, cnt);
    TAKE(changed, int, 1);
    CK(cudaMemsetAsync(v, 0, cnt * sizeof(double), ctx->stream));
    const bool vec = vec_ok(C, st, n, n);
    constexpr int V = natural_vec<CT>();
    int rounds = 0, moved = 1;
    const int max_rounds = n > 1 ? n - 1 : 1;
    while (moved && rounds < max_rounds) {
        CK(cudaMemsetAsync(changed, 0, sizeof(int), ctx->stream));
        const int burst = rounds < 8 ? 4 : 16;                  // rounds between two looks at the flag
        for (int q = 0; q < burst && rounds < max_rounds; ++q, ++rounds) {
            B200LAP_LAUNCH(k_bf_gather, dim3((n + 255) / 256, batch), dim3(256), 0, ctx->stream, (const void*)C, (int)(sizeof(CT) == 8), st, n, n, x,
                           (const double*)v, vx, cx);
            if (vec) {
                auto k = k_bf_relax_partial<CT, V>;
                B200LAP_LAUNCH(k, dim3((n + kColThreads * V - 1) / (kColThreads * V), S, batch), dim3(kColThreads), 0, ctx->stream, C, st, n, n, rps,
                               (const double*)vx, (const double*)cx, pval);
            } else {
                auto k = k_bf_relax_partial<CT, 1>;
                B200LAP_LAUNCH(k, dim3((n + kColThreads - 1) / kColThreads, S, batch), dim3(kColThreads), 0, ctx->stream, C, st, n, n, rps,
                               (const double*)vx, (const double*)cx, pval);
            }
            B200LAP_LAUNCH(k_min_trick_final, dim3((n + 255) / 256, batch), dim3(256), 0, ctx->stream, (const double*)pval, S, n, cand);
            B200LAP_LAUNCH(k_bf_update, dim3((unsigned)((cnt + 255) / 256)), dim3(256), 0, ctx->stream, v, (const double*)cand, (long long)cnt, changed);
            ctx->launches += 4;
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&moved, changed, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (rounds_out) *rounds_out = rounds;
    // still moving after n - 1 rounds: the reference raises "Negative cycle ..." (the matching was not optimal)
    return moved ? fail(B200LAP_ERR_ARG, "difference constraints did not converge in n - 1 rounds: the matching is not optimal (negative cycle)") : 0;
}

}  // namespace

extern "C" {

int b200lap_dev_project_feasible(b200lap_ctx* ctx, const void* C, int is_f64, int n, double* u, double* v, int max_rounds,
                                 double tol, int* rounds) {
    if (!ctx) return fail(B200LAP_ERR_ARG, "ctx is null");
    if (n <= 0) return fail(B200LAP_ERR_ARG, "empty problem");
    ctx->ws_reset();
    return is_f64 ? run_project_feasible(ctx, (const double*)C, n, u, v, max_rounds, tol, rounds)
                  : run_project_feasible(ctx, (const float*)C, n, u, v, max_rounds, tol, rounds);
}

int b200lap_dev_reduced_costs(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, const double* u, const double* v,
                              double* out, double* min_host) {
    if (!ctx || !min_host) return fail(B200LAP_ERR_ARG, "null argument");
    if (n <= 0 || batch <= 0) return fail(B200LAP_ERR_ARG, "empty problem");
    ctx->ws_reset();
    const long long st = (long long)n * n;
    return is_f64 ? run_reduced_costs(ctx, (const double*)C, st, n, batch, n, u, v, out, min_host)
                  : run_reduced_costs(ctx, (const float*)C, st, n, batch, n, u, v, out, min_host);
}

int b200lap_dev_bf_duals(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, const int* x, double* v, int* rounds) {
    if (!ctx || !C || !x || !v) return fail(B200LAP_ERR_ARG, "null argument");
    if (n <= 0 || batch <= 0) return fail(B200LAP_ERR_ARG, "empty problem");
    ctx->ws_reset();
    return is_f64 ? run_bf_duals(ctx, (const double*)C, batch, n, x, v, rounds) : run_bf_duals(ctx, (const float*)C, batch, n, x, v, rounds);
}

}  // extern "C"

extern "C" {

// ---- host-buffer entry points -------------------------------------------------------------------
namespace {

struct HostMatrix {
    const void* dev = nullptr;   // device matrix actually used
    int is_f64 = 0;
};

// Upload `batch` binary64 matrices and, when every entry survives the binary32 round trip,
// switch to a binary32 device copy (half the HBM traffic for every later sweep).
int upload_matrices(b200lap_ctx* ctx, const double* C, int batch, int n, HostMatrix* out) {
    const size_t count = (size_t)batch * n * n;
    TAKE(d64, double, count);
    TAKE(d32, float, count);
    TAKE(dflag, int, 1);
    CK(cudaMemcpyAsync(d64, C, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int one = 1;
    CK(cudaMemcpyAsync(dflag, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    int r = run_narrow(ctx, d64, (long long)count, d32, dflag);
    if (r) return r;
    int exact = 0;
    CK(cudaMemcpyAsync(&exact, dflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out->is_f64 = exact ? 0 : 1;
    out->dev = exact ? (const void*)d32 : (const void*)d64;
    return 0;
}

int solve_host(b200lap_ctx* ctx, const double* C, int batch, int n, long long* x, long long* y, const double* u_seed,
               const double* v_seed, double eps, int mode, int* rc_out, long long* trace_out)
{
    std::lock_guard<std::mutex> g(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    ctx->ws_reset();
    HostMatrix M;
    int r = upload_matrices(ctx, C, batch, n, &M);
    if (r) return r;
    const size_t bn = (size_t)batch * n;
    double *du = nullptr, *dv = nullptr;
    if (mode == 0) {
        du = ctx->take_n<double>(bn);
        dv = ctx->take_n<double>(bn);
        if (!du || !dv) return fail(-1, "device workspace allocation failed (seeds)");
        CK(cudaMemcpyAsync(du, u_seed, bn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(dv, v_seed, bn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    TAKE(dx, int, bn);
    TAKE(dy, int, bn);
    TAKE(drc, int, (size_t)batch);
    long long* dtr = nullptr;
    if (trace_out) {
        dtr = ctx->take_n<long long>((size_t)batch * kTraceWords);
        if (!dtr) return fail(-1, "device workspace allocation failed (trace)");
    }
    const long long st = (long long)n * n;
    r = M.is_f64 ? run_solve(ctx, (const double*)M.dev, st, n, batch, n, du, dv, eps, mode, dx, dy, drc, dtr, nullptr)
                 : run_solve(ctx, (const float*)M.dev, st, n, batch, n, du, dv, eps, mode, dx, dy, drc, dtr, nullptr);
    if (r) return r;
    std::vector<int> hx(bn), hy(bn), hrc((size_t)batch);
    CK(cudaMemcpyAsync(hx.data(), dx, bn * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hy.data(), dy, bn * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hrc.data(), drc, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (trace_out) CK(cudaMemcpyAsync(trace_out, dtr, (size_t)batch * kTraceWords * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < batch; ++b) {
        rc_out[b] = hrc[(size_t)b];
        if (hrc[(size_t)b] != 0) continue;
        for (int k = 0; k < n; ++k) {
            x[(size_t)b * n + k] = hx[(size_t)b * n + k];
            y[(size_t)b * n + k] = hy[(size_t)b * n + k];
        }
    }
    return 0;
}

}  // namespace

int lapjv_seeded(const double* C, int n_rows, int n_cols, long long* x, long long* y, const double* u_seed,
                 const double* v_seed, double eps) {
    if (n_rows <= 0 || n_cols <= 0) return -2;
    if (n_rows != n_cols) return -4;
    try {
        b200lap_ctx* ctx = b200lap_default_ctx();
        if (!ctx) return B200LAP_ERR_CUDA;
        int rc = 0;
        const int r = solve_host(ctx, C, 1, n_rows, x, y, u_seed, v_seed, eps, 0, &rc, nullptr);
        return r ? r : rc;
    } catch (const std::bad_alloc&) {
        return -1;
    }
}

int b200lap_lapjv_seeded_batch(const double* C, int batch, int n, long long* x, long long* y, const double* u_seed,
                               const double* v_seed, double eps, int* rc, long long* trace) {
    if (batch <= 0) return 0;
    if (n <= 0) return -2;
    try {
        b200lap_ctx* ctx = b200lap_default_ctx();
        if (!ctx) return B200LAP_ERR_CUDA;
        return solve_host(ctx, C, batch, n, x, y, u_seed, v_seed, eps, 0, rc, trace);
    } catch (const std::bad_alloc&) {
        return -1;
    }
}

int b200lap_lapjv(const double* C, int n, int* x, int* y) {
    if (n <= 0) return -2;
    try {
        b200lap_ctx* ctx = b200lap_default_ctx();
        if (!ctx) return B200LAP_ERR_CUDA;
        std::vector<long long> lx((size_t)n), ly((size_t)n);
        int rc = 0;
        const int r = solve_host(ctx, C, 1, n, lx.data(), ly.data(), nullptr, nullptr, 0.0, 1, &rc, nullptr);
        if (r) return r;
        if (rc) return rc;
        for (int k = 0; k < n; ++k) { x[k] = (int)lx[(size_t)k]; y[k] = (int)ly[(size_t)k]; }
        return 0;
    } catch (const std::bad_alloc&) {
        return -1;
    }
}

}  // extern "C"

#include "api_dense.inc"
