// tests/emul/cuda_emul.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny SIMT interpreter that lets g++ compile the kernels under
// gnn-accelerated-lap-warm-start-pipeline_b200/csrc/ and run them on the CPU, one thread block at a
// time, every CUDA thread a ucontext fiber.  It exists so the bit-exact tie-breaking logic of
// the solver kernels (and the host orchestration in api.cu) can be exercised in the GPU-less
// authoring container under `pytest -m "not gpu"`, with printf/ASan/gdb available.  It is NOT
// a product code path: the shipped library is built by nvcc only (see build.py), nothing under
// the package imports or links this file, and the product fails loudly without a CUDA device.
//
// Supported subset: __syncthreads/__syncwarp, full-warp shuffles/ballots/reductions/match,
// atomics on shared/global words, static and dynamic shared memory, 1-D/3-D grids, and a shim
// of the handful of CUDA runtime calls api.cu makes (malloc/free/memcpy/memset/streams as
// synchronous host operations).  Blocks run sequentially; blockDim.x must be a multiple of 32.
#pragma once
#include <setjmp.h>
#include <ucontext.h>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define B200LAP_EMUL 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __align__(x) __attribute__((aligned(x)))
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }

namespace emul {

constexpr size_t kStack = 256 * 1024;

// Fibers are created with makecontext and switched with _setjmp / _longjmp afterwards: swapcontext saves and restores
// the signal mask with a system call on every switch, which dominated the run time of the CPU suite.
struct Fiber {
    ucontext_t ctx;
    jmp_buf jb;
    bool started = false;
    unsigned char* stack = nullptr;      // from a process-wide pool, reused by every block (never zero-filled)
    bool done = false;
    uint3 tid{0, 0, 0};
};

struct WarpSlot {
    unsigned long long val[32];
    int arrived = 0;
    int gen = 0;
};

struct Block {
    std::vector<Fiber> fibers;
    std::vector<WarpSlot> warps;
    ucontext_t sched;
    jmp_buf sched_jb;
    int cur = -1;
    int bar_arrived = 0;
    int bar_gen = 0;
    int named_arrived[16] = {0};
    int named_gen[16] = {0};
    int nthreads = 0;
    const std::function<void()>* body = nullptr;
};

inline Block*& g_block() { static Block* b = nullptr; return b; }
inline dim3& g_blockDim() { static dim3 d; return d; }
inline dim3& g_gridDim() { static dim3 d; return d; }
inline uint3& g_blockIdx() { static uint3 d{0, 0, 0}; return d; }
inline std::vector<unsigned char>& g_dyn() { static std::vector<unsigned char> v; return v; }

inline void yield() {
    Block* b = g_block();
    if (_setjmp(b->fibers[b->cur].jb) == 0) _longjmp(b->sched_jb, 1);
}

inline void block_barrier() {
    Block* b = g_block();
    int g = b->bar_gen;
    if (++b->bar_arrived == b->nthreads) { b->bar_arrived = 0; b->bar_gen++; return; }
    while (b->bar_gen == g) yield();
}

// bar.sync id, count: the first `count` threads to arrive at barrier `id` release each other
inline void named_barrier(int id, int count) {
    Block* b = g_block();
    int g = b->named_gen[id];
    if (++b->named_arrived[id] == count) { b->named_arrived[id] = 0; b->named_gen[id]++; return; }
    while (b->named_gen[id] == g) yield();
}

inline void warp_barrier() {
    Block* b = g_block();
    WarpSlot& w = b->warps[b->cur >> 5];
    int g = w.gen;
    if (++w.arrived == 32) { w.arrived = 0; w.gen++; return; }
    while (w.gen == g) yield();
}

inline int lane() { return g_block()->cur & 31; }

template <typename T> inline unsigned long long to_bits(T v) {
    static_assert(sizeof(T) <= 8, "");
    unsigned long long r = 0;
    std::memcpy(&r, &v, sizeof(T));
    return r;
}
template <typename T> inline T from_bits(unsigned long long r) {
    T v;
    std::memcpy(&v, &r, sizeof(T));
    return v;
}

// every lane publishes, then reads lane `src` (warp converged, full mask)
template <typename T> inline T exchange(T v, int src) {
    Block* b = g_block();
    WarpSlot& w = b->warps[b->cur >> 5];
    w.val[lane()] = to_bits(v);
    warp_barrier();
    T r = from_bits<T>(w.val[src & 31]);
    warp_barrier();
    return r;
}

inline void fiber_main() {
    Block* b = g_block();
    (*b->body)();
    b->fibers[b->cur].done = true;
    _longjmp(b->sched_jb, 1);
}

inline void run_block(const std::function<void()>& body, dim3 bd) {
    Block blk;
    int T = (int)(bd.x * bd.y * bd.z);
    assert(T % 32 == 0 && "emulated blocks must be whole warps");
    blk.nthreads = T;
    blk.fibers.resize(T);
    blk.warps.resize(T / 32);
    blk.body = &body;
    g_block() = &blk;
    for (int t = 0; t < T; ++t) {
        Fiber& f = blk.fibers[t];
        static std::vector<unsigned char*> pool;
        while ((int)pool.size() <= t) pool.push_back((unsigned char*)std::malloc(kStack));
        f.stack = pool[t];
        f.tid = uint3{(unsigned)(t % bd.x), (unsigned)((t / bd.x) % bd.y), (unsigned)(t / (bd.x * bd.y))};
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &blk.sched;
        makecontext(&f.ctx, (void (*)())fiber_main, 0);
    }
    int live = T;
    while (live > 0) {
        live = 0;
        for (int t = 0; t < T; ++t) {
            if (blk.fibers[t].done) continue;
            blk.cur = t;
            if (_setjmp(blk.sched_jb) == 0) {
                Fiber& f = blk.fibers[t];
                if (!f.started) { f.started = true; setcontext(&f.ctx); }
                _longjmp(f.jb, 1);
            }
            if (!blk.fibers[t].done) ++live;
        }
    }
    g_block() = nullptr;
}

inline void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body) {
    g_blockDim() = block;
    g_gridDim() = grid;
    g_dyn().assign(dyn_smem + 64, 0);
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                g_blockIdx() = uint3{x, y, z};
                run_block(body, block);
            }
}

inline unsigned char* dyn_smem_ptr() {
    uintptr_t p = (uintptr_t)g_dyn().data();
    p = (p + 63) & ~(uintptr_t)63;
    return (unsigned char*)p;
}

}  // namespace emul

#define threadIdx (emul::g_block()->fibers[emul::g_block()->cur].tid)
#define blockIdx (emul::g_blockIdx())
#define blockDim (emul::g_blockDim())
#define gridDim (emul::g_gridDim())

static inline void __syncthreads() { emul::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emul::exchange(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emul::exchange(v, emul::lane() ^ m); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int l = emul::lane();
    T r = emul::exchange(v, l >= (int)d ? l - (int)d : l);
    return l >= (int)d ? r : v;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    int l = emul::lane();
    T r = emul::exchange(v, l + (int)d < 32 ? l + (int)d : l);
    return l + (int)d < 32 ? r : v;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    emul::Block* b = emul::g_block();
    emul::WarpSlot& w = b->warps[b->cur >> 5];
    w.val[emul::lane()] = pred ? 1ull : 0ull;
    emul::warp_barrier();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(w.val[l] & 1ull) << l;
    emul::warp_barrier();
    return r;
}
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0u; }
template <typename T, typename F> static inline T emul_warp_fold(T v, F f) {
    emul::Block* b = emul::g_block();
    emul::WarpSlot& w = b->warps[b->cur >> 5];
    w.val[emul::lane()] = emul::to_bits(v);
    emul::warp_barrier();
    T r = emul::from_bits<T>(w.val[0]);
    for (int l = 1; l < 32; ++l) r = f(r, emul::from_bits<T>(w.val[l]));
    emul::warp_barrier();
    return r;
}
static inline int __reduce_add_sync(unsigned, int v) { return emul_warp_fold(v, [](int a, int b) { return a + b; }); }
static inline int __reduce_min_sync(unsigned, int v) { return emul_warp_fold(v, [](int a, int b) { return a < b ? a : b; }); }
static inline int __reduce_max_sync(unsigned, int v) { return emul_warp_fold(v, [](int a, int b) { return a > b ? a : b; }); }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return emul_warp_fold(v, [](unsigned a, unsigned b) { return a + b; }); }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return emul_warp_fold(v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return emul_warp_fold(v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return emul_warp_fold(v, [](unsigned a, unsigned b) { return a | b; }); }
template <typename T> static inline unsigned __match_any_sync(unsigned, T v) {
    emul::Block* b = emul::g_block();
    emul::WarpSlot& w = b->warps[b->cur >> 5];
    w.val[emul::lane()] = emul::to_bits(v);
    emul::warp_barrier();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(w.val[l] == emul::to_bits(v)) << l;
    emul::warp_barrier();
    return r;
}

// atomics: fibers are cooperative, so plain read-modify-write is atomic
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; *p = v < o ? v : o; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; *p = v > o ? v : o; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __float_as_uint(float f) { return emul::from_bits<unsigned>(emul::to_bits(f)); }
static inline float __uint_as_float(unsigned u) { return emul::from_bits<float>((unsigned long long)u); }
static inline int __float_as_int(float f) { return (int)__float_as_uint(f); }
static inline float __int_as_float(int i) { return __uint_as_float((unsigned)i); }
static inline long long __double_as_longlong(double d) { return (long long)emul::to_bits(d); }
static inline double __longlong_as_double(long long l) { return emul::from_bits<double>((unsigned long long)l); }
#define __expf(x) (std::exp((float)(x)))
#define __logf(x) (std::log((float)(x)))
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
using std::min;
using std::max;

// ---- CUDA runtime shim (synchronous host operations) ------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
#define cudaEventDisableTiming 2
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* c) { *c = 1; return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
#define cudaStreamNonBlocking 1

#define B200LAP_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emul::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define B200LAP_DYN_SMEM(name) unsigned char* name = emul::dyn_smem_ptr()
