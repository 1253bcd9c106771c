// host_narrow.cpp -- host-side marshalling for the host-buffer entry points: binary64 -> binary32 with an exactness
// check, multi-threaded.
//
// The reference's API hands over binary64 matrices (LAP/lap/_seeded_jv.pyx:14-31, scripts/gnn_benchmark.py:226).  The
// benchmark families are binary32-representable, and the device kernels then work on the binary32 copy anyway
// (common.cuh), so uploading 8 bytes per entry only to narrow them on the device makes PCIe the ceiling of the
// end-to-end rate (2 GiB per 64 x 2048 batch = 39 ms, 1640 instances/s).  Narrowing on the host while the previous
// batches are being solved halves the upload.  This is data marshalling, not a compute fallback: when any entry does
// not survive the round trip the caller uploads the binary64 matrix as before.
#include <atomic>
#include <cstddef>
#include <cstdlib>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace b200lap_host {

static bool narrow_scalar(const double* src, float* dst, size_t n) {
    bool ok = true;
    for (size_t i = 0; i < n; ++i) {
        const float f = (float)src[i];
        dst[i] = f;
        ok &= ((double)f == src[i]);
    }
    return ok;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static bool narrow_avx2(const double* src, float* dst, size_t n) {
    __m256d bad = _mm256_setzero_pd();
    size_t i = 0;
    const bool aligned = (((size_t)dst) & 31) == 0;
    for (; i + 8 <= n; i += 8) {
        const __m256d a = _mm256_loadu_pd(src + i), b = _mm256_loadu_pd(src + i + 4);
        const __m128 fa = _mm256_cvtpd_ps(a), fb = _mm256_cvtpd_ps(b);
        const __m256 f = _mm256_insertf128_ps(_mm256_castps128_ps256(fa), fb, 1);
        if (aligned) _mm256_stream_ps(dst + i, f); else _mm256_storeu_ps(dst + i, f);     // the staging buffer is read by DMA, not by us
        bad = _mm256_or_pd(bad, _mm256_or_pd(_mm256_cmp_pd(a, _mm256_cvtps_pd(fa), _CMP_NEQ_UQ), _mm256_cmp_pd(b, _mm256_cvtps_pd(fb), _CMP_NEQ_UQ)));
    }
    _mm_sfence();
    bool ok = _mm256_movemask_pd(bad) == 0;
    if (i < n) ok &= narrow_scalar(src + i, dst + i, n - i);
    return ok;
}
#endif

// threads the narrowing may use: B200LAP_HOST_NARROW_THREADS (0 = off), default min(8, hardware threads)
int narrow_threads() {
    static const int t = [] {
        if (const char* e = std::getenv("B200LAP_HOST_NARROW_THREADS")) return std::atoi(e) < 0 ? 0 : std::atoi(e);
        const unsigned hw = std::thread::hardware_concurrency();
        return (int)(hw >= 8 ? 8 : (hw > 1 ? hw : 1));
    }();
    return t;
}

// share (percent) of a batch's instances narrowed on the host; the rest is uploaded as binary64 by DMA meanwhile
int narrow_percent() {
    static const int p = [] {
        int v = 70;
        if (const char* e = std::getenv("B200LAP_HOST_NARROW_PERCENT")) v = std::atoi(e);
        return v < 0 ? 0 : (v > 100 ? 100 : v);
    }();
    return p;
}

// dst[i] = (float)src[i]; returns true when every entry survived the round trip (NaNs do not)
bool narrow(const double* src, float* dst, size_t count, int threads) {
    if (threads < 1) threads = 1;
    const size_t per = ((count / (size_t)threads) + 7) & ~(size_t)7;
    std::atomic<bool> ok{true};
    auto work = [&](size_t lo, size_t hi) {
        if (lo >= hi) return;
#if defined(__x86_64__)
        const bool r = __builtin_cpu_supports("avx2") ? narrow_avx2(src + lo, dst + lo, hi - lo) : narrow_scalar(src + lo, dst + lo, hi - lo);
#else
        const bool r = narrow_scalar(src + lo, dst + lo, hi - lo);
#endif
        if (!r) ok.store(false, std::memory_order_relaxed);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) {
        const size_t lo = (size_t)t * per, hi = t == threads - 1 ? count : (size_t)(t + 1) * per;
        if (lo < count) pool.emplace_back(work, lo, hi < count ? hi : count);
    }
    work(0, per < count ? per : count);
    for (auto& th : pool) th.join();
    return ok.load();
}

}  // namespace b200lap_host
