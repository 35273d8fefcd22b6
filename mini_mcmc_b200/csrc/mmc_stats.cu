// K5: split-Rhat / ESS on device (see include/minimcmc.h "diagnostics").
// Reproduces split_rhat_mean_ess, src/stats.rs:396-554:
//   split every chain into its first and LAST n/2 draws -> C = 2c split chains of N = n/2 draws;
//   per parameter: chain means m_j, biased variances s_j, W = mean s_j, B = N/(C-1) sum (m_j - mbar)^2,
//   var+ = (N-1)/N W + B/N, rhat = sqrt(W / var+) (sic), autocovariance acov_j(t) (1/N normalisation),
//   rho_t = 1 - (W - mean_j acov_j(t)) / var+, Geyer initial-positive/monotone pair sums, ESS = C N / tau.
//
// Device part: ONE streaming pass over the draws per block of 16 lags.  A thread owns one
// (split chain, parameter) series; the 32 lanes of a warp own 32 adjacent parameters so every load is a
// contiguous 128 B row segment; the last 16 values live in a register ring so each draw is read once.
// Cross-chain sums are accumulated in f64 registers by a persistent
// grid and flushed with one f64 atomic per (CTA, parameter, row).
// Only lags the Geyer truncation actually consumes are computed (the reference computes all N by FFT
// and then discards everything after the first non-positive pair); results differ from the FFT path
// by f32 rounding only.
//
// Host part (mmc_stats_finalize): the O(p * lags) Geyer loop and basic_stats (src/stats.rs:310-336).
#include <algorithm>
#include <cmath>
#include <vector>

#include "mmc_common.cuh"

using namespace mmc;

namespace {

constexpr int kLagBlock = 16;
constexpr int kStatsWarps = 4;

// partial layout: [2 + N][p] doubles: row 0 = sum_j m_j, row 1 = sum_j m_j^2, row 2 + t = sum_j acov_j(t)
//
// Per series: loop 1 sums the draws (mean, src/stats.rs:586), loop 2 re-reads them (L2 hits: the chain block was
// just streamed) and accumulates sum_t d_t d_{t-lag} for 16 lags with the last 16 centred values in a register
// ring.  All loads are unconditional (indices clamped, contributions masked) so 16 are in flight per thread.
template <bool kFirst>
__global__ void __launch_bounds__(kStatsWarps * 32, 4)
stats_pass_kernel(const float *__restrict__ sample, int64_t c_local, int64_t n, int64_t p, int64_t lag0, int nlag,
                  double *__restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * kStatsWarps + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * kStatsWarps;
    const int64_t nb = (p + 31) / 32;            // parameter blocks
    const int64_t b = warp_global % nb;          // this warp's parameter block (fixed for its lifetime)
    const int64_t q = b * 32 + lane;
    const bool q_ok = q < p;
    const int64_t qc = q_ok ? q : p - 1;         // out-of-range lanes shadow the last parameter (never flushed)
    const int64_t half = n / 2, N = half, C = 2 * c_local;
    const int64_t chain_stride = n_warps / nb;   // host guarantees n_warps % nb == 0
    const float inv_n = 1.0f / (float)N;

    double acc_m = 0.0, acc_m2 = 0.0, acc_cov[kLagBlock];
#pragma unroll
    for (int i = 0; i < kLagBlock; ++i) acc_cov[i] = 0.0;

    for (int64_t j = warp_global / nb; j < C; j += chain_stride) {
        // splitcat, src/stats.rs:396-402: split chain j < c is the first half of chain j, j >= c the last half of j - c
        const int64_t chain = j < c_local ? j : j - c_local;
        const int64_t row0 = j < c_local ? 0 : n - half;
        const float *base = sample + ((chain * n + row0) * p + qc);
        // ---- loop 1: mean
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
        int64_t t0 = 0;
        for (; t0 + kLagBlock <= N; t0 += kLagBlock) {
            float v[kLagBlock];
#pragma unroll
            for (int u = 0; u < kLagBlock; ++u) v[u] = __ldg(base + (t0 + u) * p);
#pragma unroll
            for (int u = 0; u < kLagBlock; u += 4) { s0 += v[u]; s1 += v[u + 1]; s2 += v[u + 2]; s3 += v[u + 3]; }
        }
        for (; t0 < N; ++t0) s0 += __ldg(base + t0 * p);
        const float m = ((s0 + s1) + (s2 + s3)) * inv_n;
        // ---- loop 2: centred lagged products, lags lag0 .. lag0 + 15
        float P[kLagBlock], ring[kLagBlock];
#pragma unroll
        for (int i = 0; i < kLagBlock; ++i) { P[i] = 0.0f; ring[i] = 0.0f; }
        for (t0 = 0; t0 < N; t0 += kLagBlock) {
            float a[kLagBlock], bb[kLagBlock];
#pragma unroll
            for (int u = 0; u < kLagBlock; ++u) {
                const int64_t t = t0 + u;
                const int64_t tc = t < N ? t : N - 1;
                const float va = __ldg(base + tc * p);
                a[u] = t < N ? va - m : 0.0f;
                if (kFirst) {
                    bb[u] = a[u];
                } else {
                    const int64_t tb = tc - lag0;
                    const float vb = __ldg(base + (tb > 0 ? tb : 0) * p);
                    bb[u] = (t < N && tb >= 0) ? vb - m : 0.0f;
                }
            }
#pragma unroll
            for (int u = 0; u < kLagBlock; ++u) {
                ring[u] = bb[u];  // b_t at slot t mod 16
#pragma unroll
                for (int i = 0; i < kLagBlock; ++i) P[i] = fmaf(a[u], ring[(u - i) & (kLagBlock - 1)], P[i]);
            }
        }
        if (kFirst) {
            acc_m += (double)m;
            acc_m2 += (double)m * (double)m;
        }
#pragma unroll
        for (int i = 0; i < kLagBlock; ++i) acc_cov[i] += (double)(P[i] * inv_n);
    }
    if (q_ok) {
        if (kFirst) {
            atomicAdd(partial + q, acc_m);
            atomicAdd(partial + p + q, acc_m2);
        }
#pragma unroll
        for (int i = 0; i < kLagBlock; ++i)
            if (i < nlag && lag0 + i < N) atomicAdd(partial + (2 + lag0 + i) * p + q, acc_cov[i]);
    }
}

int launch_pass(const float *sample, int64_t c_local, int64_t n, int64_t p, int64_t lag0, int nlag, double *partial,
                cudaStream_t stream) {
    const int64_t nb = (p + 31) / 32;
    // persistent grid: ~8 CTAs per SM, rounded so that the warp count is a multiple of the parameter blocks
    int64_t warps = (int64_t)sm_count() * 4 * kStatsWarps;
    const int64_t work = 2 * c_local * nb;
    if (warps > work) warps = work;
    warps = ((warps + nb * kStatsWarps - 1) / (nb * kStatsWarps)) * (nb * kStatsWarps);
    const unsigned grid = (unsigned)(warps / kStatsWarps);
    if (lag0 == 0)
        stats_pass_kernel<true><<<grid, kStatsWarps * 32, 0, stream>>>(sample, c_local, n, p, lag0, nlag, partial);
    else
        stats_pass_kernel<false><<<grid, kStatsWarps * 32, 0, stream>>>(sample, c_local, n, p, lag0, nlag, partial);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

}  // namespace

extern "C" {

int64_t mmc_stats_partial_len(int64_t n, int64_t p) { return (2 + n / 2) * p; }

int mmc_stats_partial_dev(const float *sample_dev, int64_t c_local, int64_t n, int64_t p, int64_t lag0,
                          int64_t n_lags, double *partial_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(sample_dev && partial_dev && c_local > 0 && n >= 2 && p > 0 && lag0 >= 0 && n_lags > 0,
                "mmc_stats_partial_dev: bad arguments");
    const int64_t N = n / 2;
    MMC_REQUIRE(lag0 < N, "lag0 %lld >= N %lld", (long long)lag0, (long long)N);
    if (lag0 + n_lags > N) n_lags = N - lag0;
    cudaStream_t s = (cudaStream_t)stream;
    // zero the rows this call produces
    if (lag0 == 0) MMC_CUDA(cudaMemsetAsync(partial_dev, 0, sizeof(double) * 2 * p, s));
    MMC_CUDA(cudaMemsetAsync(partial_dev + (2 + lag0) * p, 0, sizeof(double) * n_lags * p, s));
    for (int64_t l = 0; l < n_lags; l += kLagBlock) {
        const int nl = (int)std::min<int64_t>(kLagBlock, n_lags - l);
        rc = launch_pass(sample_dev, c_local, n, p, lag0 + l, nl, partial_dev, s);
        if (rc) return rc;
    }
    return MMC_OK;
}

// returns 0 when every parameter's Geyer sum terminated inside the available lags (or all N lags are
// available), 1 when more lags are needed.
int mmc_stats_finalize(const double *partial, int64_t c_total, int64_t n, int64_t p, int64_t lags_available,
                       float *rhat_out, float *ess_out) {
    MMC_REQUIRE(partial && c_total > 0 && n >= 2 && p > 0 && lags_available > 0, "mmc_stats_finalize: bad arguments");
    const int64_t N = n / 2;
    const double C = 2.0 * (double)c_total;
    if (lags_available > N) lags_available = N;
    int need_more = 0;
    for (int64_t q = 0; q < p; ++q) {
        const double sm = partial[q], sm2 = partial[p + q];
        const double mbar = sm / C;
        double ssd = sm2 - C * mbar * mbar;  // sum_j (m_j - mbar)^2
        if (ssd < 0.0) ssd = 0.0;
        const double B = ssd * ((double)N / (C - 1.0));
        const double W = partial[2 * p + q] / C;  // mean_j s_j  (s_j = acov_j(0))
        const double var = (((double)N - 1.0) / (double)N) * W + B / (double)N;
        const float within = (float)W, varf = (float)var;
        if (rhat_out) rhat_out[q] = sqrtf(within / varf);  // src/stats.rs:425-427 (sic)
        // Geyer, src/stats.rs:518-545, in f32 like the reference
        auto rho = [&](int64_t t) {
            const float avg = (float)(partial[(2 + t) * p + q] / C);
            return -((-avg + within) / varf) + 1.0f;
        };
        float mn = N >= 2 ? rho(0) + rho(1) : 0.0f;
        float o = 0.0f;
        bool terminated = false;
        int64_t t = 0;
        for (; t + 1 < lags_available; t += 2) {
            float pt = rho(t) + rho(t + 1);
            if (pt <= 0.0f) { terminated = true; break; }
            if (pt > mn) pt = mn;
            mn = pt;
            o += pt;
        }
        if (!terminated && t + 1 < N) need_more = 1;  // ran out of computed lags before the window ended
        const float tau = -1.0f + 2.0f * o;
        if (ess_out) ess_out[q] = (1.0f / tau) * (float)C * (float)N;
    }
    return need_more;
}

int mmc_split_rhat_ess_dev(const float *sample_dev, int64_t c, int64_t n, int64_t p, float *rhat_host,
                           float *ess_host, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(sample_dev && c > 0 && n >= 2 && p > 0, "mmc_split_rhat_ess_dev: bad arguments");
    const int64_t N = n / 2;
    const int64_t len = mmc_stats_partial_len(n, p);
    cudaStream_t s = (cudaStream_t)stream;
    double *d_partial = nullptr;
    MMC_CUDA(cudaMalloc((void **)&d_partial, sizeof(double) * len));
    std::vector<double> h_partial((size_t)len, 0.0);
    int64_t have = 0, block = kLagBlock;
    int result = MMC_OK;
    while (have < N) {
        const int64_t want = std::min<int64_t>(block, N - have);
        rc = mmc_stats_partial_dev(sample_dev, c, n, p, have, want, d_partial, stream);
        if (rc) { result = rc; break; }
        const size_t off = have == 0 ? 0 : (size_t)(2 + have) * p;
        const size_t cnt = (have == 0 ? 2 * p : 0) + (size_t)want * p;
        cudaError_t e = cudaMemcpyAsync(h_partial.data() + off, d_partial + off, cnt * sizeof(double),
                                        cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { result = cuda_fail(e, "stats D2H", __FILE__, __LINE__); break; }
        have += want;
        const int more = mmc_stats_finalize(h_partial.data(), c, n, p, have, rhat_host, ess_host);
        if (more < 0) { result = more; break; }
        if (more == 0) break;
        block *= 2;  // geometric growth keeps the number of host round trips logarithmic
    }
    cudaFree(d_partial);
    return result;
}

int mmc_split_rhat_ess(const float *sample_host, int64_t c, int64_t n, int64_t p, float *rhat_host, float *ess_host) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(sample_host && c > 0 && n >= 2 && p > 0, "mmc_split_rhat_ess: bad arguments");
    float *d = nullptr;
    const size_t bytes = sizeof(float) * (size_t)c * n * p;
    MMC_CUDA(cudaMalloc((void **)&d, bytes));
    cudaError_t e = cudaMemcpy(d, sample_host, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "stats H2D", __FILE__, __LINE__); }
    rc = mmc_split_rhat_ess_dev(d, c, n, p, rhat_host, ess_host, nullptr);
    cudaFree(d);
    return rc;
}

// basic_stats, src/stats.rs:310-336: sort descending; min = last, max = first, median = data[len/2],
// mean, std with ddof = 1.
int mmc_basic_stats_of(const float *data, int64_t len, mmc_basic_stats *out) {
    MMC_REQUIRE(data && out && len > 0, "mmc_basic_stats_of: bad arguments");
    std::vector<float> d(data, data + len);
    std::stable_sort(d.begin(), d.end(), [](float a, float b) { return b < a; });
    float sum = 0.0f;
    for (float v : d) sum += v;
    float m = 0.0f, s2 = 0.0f;
    for (int64_t i = 0; i < len; ++i) {
        const float delta = d[i] - m;
        m += delta / (float)(i + 1);
        s2 += delta * (d[i] - m);
    }
    out->min = d[len - 1];
    out->median = d[len / 2];
    out->max = d[0];
    out->mean = sum / (float)len;
    out->std = sqrtf(s2 / ((float)len - 1.0f));
    return MMC_OK;
}

}  // extern "C"
