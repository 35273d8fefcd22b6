// Host half of the compact device->host path for integer MH draws (mmc_mh_run, Poisson target).
// The state of the Poisson chain fits 8 or 16 bits, but the reference API returns `usize` (u64) samples.  Moving
// u64 over PCIe costs 8 B per draw and is the end-to-end bound; instead the kernel emits u8/u16, the copy engine
// moves 1-2 B per draw, and these threads widen into the caller's [chains, n_collect] u64 array with streaming
// stores while the next block of chains is being sampled and copied.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace mmc {

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void widen_u8_avx2(const uint8_t *src, uint64_t *dst, size_t n) {
    size_t i = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    if (aligned) {
        for (; i + 16 <= n; i += 16) {
            const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
            _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_cvtepu8_epi64(v));
            _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 4), _mm256_cvtepu8_epi64(_mm_srli_si128(v, 4)));
            _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 8), _mm256_cvtepu8_epi64(_mm_srli_si128(v, 8)));
            _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 12), _mm256_cvtepu8_epi64(_mm_srli_si128(v, 12)));
        }
    }
    for (; i < n; ++i) dst[i] = src[i];
}
__attribute__((target("avx2"))) static void widen_u16_avx2(const uint16_t *src, uint64_t *dst, size_t n) {
    size_t i = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    if (aligned) {
        for (; i + 8 <= n; i += 8) {
            const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
            _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_cvtepu16_epi64(v));
            _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 4), _mm256_cvtepu16_epi64(_mm_srli_si128(v, 8)));
        }
    }
    for (; i < n; ++i) dst[i] = src[i];
}
#endif

#if defined(__x86_64__)
__attribute__((target("avx2"))) void fill_stream_avx2(uint64_t *dst, size_t n, uint64_t v) {
    const __m256i x = _mm256_set1_epi64x((long long)v);
    size_t i = 0;
    for (; i + 4 <= n; i += 4) _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), x);
    for (; i < n; ++i) dst[i] = v;
    _mm_sfence();
}
#endif

static void widen_block(const void *src, int elem_bytes, uint64_t *dst, size_t n) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) {
        if (elem_bytes == 1) widen_u8_avx2(static_cast<const uint8_t *>(src), dst, n);
        else widen_u16_avx2(static_cast<const uint16_t *>(src), dst, n);
        return;
    }
#endif
    if (elem_bytes == 1) { const uint8_t *s = static_cast<const uint8_t *>(src); for (size_t i = 0; i < n; ++i) dst[i] = s[i]; }
    else { const uint16_t *s = static_cast<const uint16_t *>(src); for (size_t i = 0; i < n; ++i) dst[i] = s[i]; }
}

int widen_threads() {
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    int ranks = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));  // one process per GPU shares the host cores
    if (const char *e = getenv("MMC_HOST_THREADS")) return std::max(1, atoi(e));
    return std::max(1, hw / ranks);
}

// dst[i] = src[i] for n elements of elem_bytes (1 or 2), split over the host threads
void widen_to_u64(const void *src, int elem_bytes, uint64_t *dst, size_t n) {
    const int nt = widen_threads();
    const size_t chunk = ((n + nt - 1) / nt + 63) & ~size_t(63);
    if (nt == 1 || n < (1u << 16)) { widen_block(src, elem_bytes, dst, n); return; }
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t) {
        const size_t b = (size_t)t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        th.emplace_back([=]() { widen_block(static_cast<const uint8_t *>(src) + b * elem_bytes, elem_bytes, dst + b, e - b); });
    }
    for (auto &x : th) x.join();
#if defined(__x86_64__)
    _mm_sfence();
#endif
}

}  // namespace mmc

// STREAM-style write peak of this host: `threads` threads (0 = the widening pool's size) fill a fresh buffer of `bytes`
// with the same non-temporal 256-bit stores the widening uses; best of `reps` passes.  bench.py reports the end-to-end
// Poisson number against it (the reference API returns u64 draws, so 8 B per draw must be written by the host cores).
extern "C" int mmc_host_write_bandwidth(uint64_t bytes, int32_t threads, int32_t reps, double *gb_per_s, int32_t *threads_used) {
    if (!gb_per_s || bytes < (1u << 20)) return -1;
    const int nt = threads > 0 ? threads : mmc::widen_threads();
    const size_t n = (size_t)bytes / 8;
    uint64_t *buf = static_cast<uint64_t *>(aligned_alloc(4096, n * 8));
    if (!buf) return -6;
    double best = 0.0;
    for (int r = 0; r < (reps > 0 ? reps : 3) + 1; ++r) {   // the first pass faults the pages in and is not timed
        const auto t0 = std::chrono::steady_clock::now();
        const size_t chunk = ((n + nt - 1) / nt + 63) & ~size_t(63);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) {
            const size_t b = (size_t)t * chunk, e = std::min(n, b + chunk);
            if (b >= e) break;
            th.emplace_back([=]() {
#if defined(__x86_64__)
                if (__builtin_cpu_supports("avx2")) { mmc::fill_stream_avx2(buf + b, e - b, (uint64_t)r); return; }
#endif
                for (size_t i = b; i < e; ++i) buf[i] = (uint64_t)r;
            });
        }
        for (auto &x : th) x.join();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (r > 0 && dt > 0.0) best = std::max(best, (double)(n * 8) / dt / 1e9);
    }
    volatile uint64_t sink = buf[n / 2];
    (void)sink;
    free(buf);
    *gb_per_s = best;
    if (threads_used) *threads_used = nt;
    return 0;
}
