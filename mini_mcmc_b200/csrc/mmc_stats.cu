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
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the symbols are resolved with dlopen (see nccl_api)

#include "mmc_common.cuh"

using namespace mmc;

extern "C" int64_t mmc_stats_partial_len(int64_t n, int64_t p);
extern "C" int mmc_stats_partial_dev(const float *sample_dev, int64_t c_local, int64_t n, int64_t p, int64_t lag0,
                                     int64_t n_lags, double *partial_dev, void *stream);

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

// ---------------------------------------------------------------- shared-memory staged variant
// When one split chain's [N, p] block fits in shared memory (C5: 200 x 100 x 4 B = 80 KB) it is fetched ONCE with a
// single bulk async copy (cp.async.bulk + mbarrier, double buffered across chains) and every lag group is computed
// from shared memory: thread (q, h) owns parameter q and the 16 lags lag0 + 16 h .. + 15, so one launch covers
// 16 * H lags while the draws cross HBM once.
__device__ __forceinline__ uint32_t st_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void st_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void st_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <bool kFirst>
__global__ void __launch_bounds__(512, 1)
stats_block_kernel(const float *__restrict__ sample, int64_t c_local, int64_t n, int p, int64_t lag0, int H, int G, int K, int nbuf,
                   double *__restrict__ partial) {
    extern __shared__ __align__(128) float st_smem[];
    __shared__ __align__(8) uint64_t bars[2];
    const int N = (int)(n / 2);
    const int64_t C = 2 * c_local;
    const int blk = N * p;                       // floats per split-chain block
    float *bufs = st_smem;                       // [nbuf][K][blk]: K split chains are staged per round (small p)
    float *mean_part = st_smem + (size_t)nbuf * K * blk;   // [H * G][K * p]
    const int tid = threadIdx.x;
    // thread (k, q, h, g): staged chain k, parameter q, lag group h (16 lags), time segment g (draws [t_lo, t_hi))
    const int PK = p * K;
    const int v = tid % PK, r = tid / PK, R = H * G;
    const int q = v % p, k = v / p;
    const int h = r % H, g = r / H;
    const bool active = r < R;
    const int seg = (((N + G - 1) / G) + kLagBlock - 1) / kLagBlock * kLagBlock;
    const int t_lo = g * seg, t_hi = (t_lo + seg < N) ? t_lo + seg : N;
    const float inv_n = 1.0f / (float)N;
    const uint32_t bytes = (uint32_t)blk * 4u;
    if (tid == 0) {
        st_mbar_init(st_smem_u32(&bars[0]), 1);
        st_mbar_init(st_smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto block_src = [&](int64_t j) {
        // splitcat, src/stats.rs:396-402
        const int64_t chain = j < c_local ? j : j - c_local;
        const int64_t row0 = j < c_local ? 0 : n - N;
        return sample + (chain * n + row0) * p;
    };
    auto issue = [&](int64_t j, int b) {
        const int nk = (int)((C - j < K) ? (C - j) : K);
        st_mbar_expect_tx(st_smem_u32(&bars[b]), bytes * (uint32_t)nk);
        for (int kk = 0; kk < nk; ++kk)
            st_bulk_load(st_smem_u32(bufs + ((size_t)b * K + kk) * blk), block_src(j + kk), bytes, st_smem_u32(&bars[b]));
    };
    float acc[kLagBlock];
#pragma unroll
    for (int i = 0; i < kLagBlock; ++i) acc[i] = 0.f;
    double acc_m = 0.0, acc_m2 = 0.0;
    const int lag_base = (int)lag0 + h * kLagBlock;

    int64_t j = (int64_t)blockIdx.x * K;
    const int64_t stride = (int64_t)gridDim.x * K;
    if (tid == 0 && j < C) issue(j, 0);
    uint32_t it = 0;
    for (; j < C; j += stride, ++it) {
        const int b = nbuf == 2 ? (int)(it & 1u) : 0;
        const uint32_t ph = nbuf == 2 ? ((it >> 1) & 1u) : (it & 1u);
        if (nbuf == 2 && tid == 0 && j + stride < C) issue(j + stride, b ^ 1);   // prefetch the next round
        st_mbar_wait(st_smem_u32(&bars[b]), ph);
        const bool have = active && (j + k < C);
        const float *x = bufs + ((size_t)b * K + k) * blk + q;
        // ---- mean: thread (k, q, r) sums t = r, r + R, ...; the R partials are combined through shared memory
        if (have) {
            float s0 = 0.f;
            for (int t = r; t < N; t += R) s0 += x[(size_t)t * p];
            mean_part[r * PK + v] = s0;
        }
        __syncthreads();
        float m = 0.f;
        if (have && t_lo < N) {
            for (int rr = 0; rr < R; ++rr) m += mean_part[rr * PK + v];
            m *= inv_n;
            // ---- centred lagged products for lags lag_base .. lag_base + 15 over this thread's time segment
            float P[kLagBlock], ring[kLagBlock];
#pragma unroll
            for (int i = 0; i < kLagBlock; ++i) P[i] = 0.f;
#pragma unroll
            for (int u = 0; u < kLagBlock; ++u) {   // warm the ring with the 16 partner values preceding the segment
                const int tb = t_lo - kLagBlock + u - lag_base;
                const float vb = x[(size_t)(tb > 0 ? tb : 0) * p];
                ring[u] = (t_lo > 0 && tb >= 0) ? vb - m : 0.f;
            }
            // full 16-blocks need no masks when the partner index cannot be negative; the rest takes the masked path
            int t0 = t_lo;
            const bool lag_zero = kFirst && h == 0;
            for (; t0 + kLagBlock <= t_hi && (lag_zero || t0 >= lag_base); t0 += kLagBlock) {
                const float *xa = x + (size_t)t0 * p;
                const float *xb = x + (size_t)(t0 - lag_base) * p;
#pragma unroll
                for (int u = 0; u < kLagBlock; ++u) {
                    const float a = xa[(size_t)u * p] - m;
                    ring[u] = lag_zero ? a : xb[(size_t)u * p] - m;
#pragma unroll
                    for (int i = 0; i < kLagBlock; ++i) P[i] = fmaf(a, ring[(u - i) & (kLagBlock - 1)], P[i]);
                }
            }
            for (; t0 < t_hi; t0 += kLagBlock) {
#pragma unroll
                for (int u = 0; u < kLagBlock; ++u) {
                    const int t = t0 + u;
                    const int tc = t < N ? t : N - 1;
                    const int tb = tc - lag_base;
                    const float va = x[(size_t)tc * p];
                    const float vb = x[(size_t)(tb > 0 ? tb : 0) * p];
                    const float a = t < N ? va - m : 0.f;
                    ring[u] = (t < N && tb >= 0) ? vb - m : 0.f;
#pragma unroll
                    for (int i = 0; i < kLagBlock; ++i) P[i] = fmaf(a, ring[(u - i) & (kLagBlock - 1)], P[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < kLagBlock; ++i) acc[i] += P[i] * inv_n;
            if (kFirst && r == 0) {
                acc_m += (double)m;
                acc_m2 += (double)m * (double)m;
            }
        }
        __syncthreads();   // everyone is done with buffer b (and mean_part) before it is refilled
        if (nbuf == 1 && tid == 0 && j + stride < C) issue(j + stride, 0);
    }
    if (active) {
        if (kFirst && r == 0) {
            atomicAdd(partial + q, acc_m);
            atomicAdd(partial + p + q, acc_m2);
        }
#pragma unroll
        for (int i = 0; i < kLagBlock; ++i)
            if (lag_base + i < N) atomicAdd(partial + (int64_t)(2 + lag_base + i) * p + q, (double)acc[i]);
    }
}

// ---- packed variant (even p): a thread owns TWO adjacent parameters and every arithmetic instruction is a packed f32x2
// operation (FFMA2 / FADD2), shared-memory reads are 64-bit: half the FFMA, LDS and address instructions per element
// (the scalar kernel executes 43 warp-instructions per element).  Used for the first pass (measured 11 % faster).
struct SF2 { unsigned long long v; };
__device__ __forceinline__ SF2 sf2_pack(float lo, float hi) { SF2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void sf2_unpack(SF2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ SF2 sf2_fma(SF2 a, SF2 b, SF2 c) { SF2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ SF2 sf2_add(SF2 a, SF2 b) { SF2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ SF2 sf2_sub(SF2 a, SF2 b) { SF2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ SF2 sf2_mul(SF2 a, SF2 b) { SF2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ SF2 sf2_ld(const float *p) { SF2 r; r.v = *reinterpret_cast<const unsigned long long *>(p); return r; }

// Thread (k, q2, r): staged chain k, parameter pair q2, replica r = (lag group h of LB lags, time segment g).  A lag group
// with base lag l only has partners for t >= l, so its G segments split [l, N) (not [0, N)): every replica of a group does
// the same number of steps.  LB = 8 for the first window (lags 0..7: 8 FFMA2 per draw keep the pass under the HBM time of
// its 80 KB block), 16 afterwards.  kZero: every replica works on lag base 0 (first window), so a step's value is its own
// partner and one shared-memory read feeds the step.  Everything a thread needs per round is computed once before the round
// loop (offsets, trip counts, the warm-up mask); a round is: mean (strided partial sums + one barrier), ring warm-up,
// unmasked blocks of LB steps, one masked block.
// kThreads x kBlocks: 512 x 1 / 384 x 1 (one CTA per SM, double-buffered staging) or 256 x 2 (two CTAs per SM, one staging
// buffer each: while one CTA waits for its bulk copy the other one computes, and their barriers do not line up).
template <int LB, bool kFirst, bool kZero, int kThreads, int kBlocks>
__global__ void __launch_bounds__(kThreads, kBlocks)
stats_block2_kernel(const float *__restrict__ sample, int64_t c_local, int64_t n, int p, int64_t lag0, int H, int G, int K, int nbuf,
                    double *__restrict__ partial) {
    extern __shared__ __align__(128) float st_smem[];
    __shared__ __align__(8) uint64_t bars[2];
    const int N = (int)(n / 2);
    const int64_t C = 2 * c_local;
    const int blk = N * p;
    const int p2 = p / 2;                           // parameter pairs
    float *bufs = st_smem;                          // [nbuf][K][blk]
    float *mean_part = st_smem + (size_t)nbuf * K * blk;   // [H * G][K * p]
    const int tid = threadIdx.x;
    const int PK = p2 * K;
    const int v = tid % PK, r = tid / PK, R = H * G;
    const int q2 = v % p2, k = v / p2;              // this thread's parameters are 2 q2 and 2 q2 + 1
    const int h = r % H, g = r / H;
    const bool active = r < R;
    const int lag_base = kZero ? 0 : (int)lag0 + h * LB;
    // this replica's steps: segment g of [lag_base, N)
    const int span = N > lag_base ? N - lag_base : 0;
    const int seg = (span + G - 1) / G;
    const int t_lo = lag_base + g * seg;
    const int t_hi = (t_lo + seg < N) ? t_lo + seg : N;
    const int n_steps = (active && t_hi > t_lo) ? t_hi - t_lo : 0;
    const int n_full = n_steps / LB, n_rem = n_steps % LB;
    const int off_a = k * blk + t_lo * p + 2 * q2;              // this thread's first draw (floats from the buffer base)
    const int off_b = off_a - lag_base * p;                     // and its partner at the group's base lag
    // ring slot u before the first block holds the partner of step t_lo - LB + u, i.e. draw t_lo - LB + u - lag_base (if any)
    uint32_t warm_mask = 0;
#pragma unroll
    for (int u = 0; u < LB; ++u)
        if (t_lo - LB + u - lag_base >= 0) warm_mask |= 1u << u;
    const int off_w = off_b - LB * p;
    // mean: replica r sums t = r, r + R, ...
    const int mean_cnt = active ? (N - r + R - 1) / R : 0;
    const int off_m = k * blk + r * p + 2 * q2, stride_m = R * p;
    const float inv_n = 1.0f / (float)N;
    const SF2 inv_n2 = sf2_pack(inv_n, inv_n), zero2 = sf2_pack(0.f, 0.f);
    const uint32_t bytes = (uint32_t)blk * 4u;
    if (tid == 0) {
        st_mbar_init(st_smem_u32(&bars[0]), 1);
        st_mbar_init(st_smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto block_src = [&](int64_t j) {
        const int64_t chain = j < c_local ? j : j - c_local;
        const int64_t row0 = j < c_local ? 0 : n - N;
        return sample + (chain * n + row0) * p;
    };
    auto issue = [&](int64_t j, int b) {
        const int nk = (int)((C - j < K) ? (C - j) : K);
        st_mbar_expect_tx(st_smem_u32(&bars[b]), bytes * (uint32_t)nk);
        for (int kk = 0; kk < nk; ++kk)
            st_bulk_load(st_smem_u32(bufs + ((size_t)b * K + kk) * blk), block_src(j + kk), bytes, st_smem_u32(&bars[b]));
    };
    SF2 acc[LB];
#pragma unroll
    for (int i = 0; i < LB; ++i) acc[i] = zero2;
    double acc_m0 = 0.0, acc_m1 = 0.0, acc_q0 = 0.0, acc_q1 = 0.0;

    int64_t j = (int64_t)blockIdx.x * K;
    const int64_t stride = (int64_t)gridDim.x * K;
    if (tid == 0 && j < C) issue(j, 0);
    uint32_t it = 0;
    for (; j < C; j += stride, ++it) {
        const int b = nbuf == 2 ? (int)(it & 1u) : 0;
        const uint32_t ph = nbuf == 2 ? ((it >> 1) & 1u) : (it & 1u);
        if (nbuf == 2 && tid == 0 && j + stride < C) issue(j + stride, b ^ 1);
        st_mbar_wait(st_smem_u32(&bars[b]), ph);
        const bool have = active && (j + k < C);
        const float *xbuf = bufs + (size_t)b * K * blk;   // 8-byte aligned offsets: blk and p are even
        if (have) {
            SF2 s0 = zero2, s1 = zero2;
            const float *pm = xbuf + off_m;
            int c2 = mean_cnt;
            for (; c2 >= 2; c2 -= 2, pm += 2 * stride_m) {
                s0 = sf2_add(s0, sf2_ld(pm));
                s1 = sf2_add(s1, sf2_ld(pm + stride_m));
            }
            if (c2) s0 = sf2_add(s0, sf2_ld(pm));
            *reinterpret_cast<unsigned long long *>(mean_part + (size_t)r * 2 * PK + 2 * v) = sf2_add(s0, s1).v;
        }
        __syncthreads();
        if (have) {
            SF2 m = zero2;
            for (int rr = 0; rr < R; ++rr) m = sf2_add(m, sf2_ld(mean_part + (size_t)rr * 2 * PK + 2 * v));
            m = sf2_mul(m, inv_n2);
            if (n_steps > 0) {
                SF2 P[LB], ring[LB];
#pragma unroll
                for (int i = 0; i < LB; ++i) P[i] = zero2;
                if (warm_mask == 0) {
#pragma unroll
                    for (int u = 0; u < LB; ++u) ring[u] = zero2;
                } else {
                    const float *pw = xbuf + off_w;
#pragma unroll
                    for (int u = 0; u < LB; ++u)
                        ring[u] = ((warm_mask >> u) & 1u) ? sf2_sub(sf2_ld(pw + u * p), m) : zero2;
                }
                const float *pa = xbuf + off_a, *pb = xbuf + off_b;
                for (int bi = 0; bi < n_full; ++bi) {
#pragma unroll
                    for (int u = 0; u < LB; ++u) {
                        const SF2 a = sf2_sub(sf2_ld(pa), m);
                        pa += p;
                        if (kZero) {
                            ring[u] = a;
                        } else {
                            ring[u] = sf2_sub(sf2_ld(pb), m);
                            pb += p;
                        }
#pragma unroll
                        for (int i = 0; i < LB; ++i) P[i] = sf2_fma(a, ring[(u - i) & (LB - 1)], P[i]);
                    }
                }
                if (n_rem) {   // last, partial block of the segment (clamped reads, masked products)
#pragma unroll
                    for (int u = 0; u < LB; ++u) {
                        const bool ok = u < n_rem;
                        const SF2 va = sf2_sub(sf2_ld(pa), m);
                        const SF2 a = ok ? va : zero2;
                        if (kZero) {
                            ring[u] = a;
                        } else {
                            ring[u] = sf2_sub(sf2_ld(pb), m);
                            if (u + 1 < n_rem) pb += p;
                        }
                        if (u + 1 < n_rem) pa += p;
#pragma unroll
                        for (int i = 0; i < LB; ++i) P[i] = sf2_fma(a, ring[(u - i) & (LB - 1)], P[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < LB; ++i) acc[i] = sf2_fma(P[i], inv_n2, acc[i]);
            }
            if (kFirst && r == 0) {
                float m0, m1;
                sf2_unpack(m, m0, m1);
                acc_m0 += (double)m0; acc_q0 += (double)m0 * (double)m0;
                acc_m1 += (double)m1; acc_q1 += (double)m1 * (double)m1;
            }
        }
        __syncthreads();
        if (nbuf == 1 && tid == 0 && j + stride < C) issue(j + stride, 0);
    }
    if (active) {
        const int q = 2 * q2;
        if (kFirst && r == 0) {
            atomicAdd(partial + q, acc_m0);
            atomicAdd(partial + q + 1, acc_m1);
            atomicAdd(partial + p + q, acc_q0);
            atomicAdd(partial + p + q + 1, acc_q1);
        }
#pragma unroll
        for (int i = 0; i < LB; ++i)
            if (lag_base + i < N) {
                float a0, a1;
                sf2_unpack(acc[i], a0, a1);
                atomicAdd(partial + (int64_t)(2 + lag_base + i) * p + q, (double)a0);
                atomicAdd(partial + (int64_t)(2 + lag_base + i) * p + q + 1, (double)a1);
            }
    }
}

// returns the number of lags covered by one launch (0 when the staged variant does not apply)
int launch_block_pass(const float *sample, int64_t c_local, int64_t n, int64_t p, int64_t lag0, int64_t n_lags,
                      double *partial, cudaStream_t stream, int64_t *covered) {
    *covered = 0;
    const int64_t N = n / 2;
    const size_t blk_bytes = (size_t)N * p * 4;
    // bulk copies need 16-byte aligned, 16-byte granular split-chain blocks (both halves of every chain)
    if (p > 512 || blk_bytes > 200 * 1024 || blk_bytes % 16 != 0 || ((size_t)n * p * 4) % 16 != 0 || ((size_t)(n - N) * p * 4) % 16 != 0 ||
        getenv("MMC_STATS_NO_SMEM"))
        return MMC_OK;
    if ((reinterpret_cast<uintptr_t>(sample) & 15) != 0) return MMC_OK;
    if (p % 2 == 0 && p <= 512 && !getenv("MMC_STATS_NO_PACKED")) {
        // packed kernel: a thread owns two adjacent parameters; up to 512 threads = (p / 2) x K staged chains x R replicas,
        // R = H lag groups x G time segments.  The first window uses groups of 8 lags, later ones 16.
        const int64_t p2 = p / 2;
        const bool first = lag0 == 0;
        const int LB = (first && n_lags <= 8) ? 8 : kLagBlock;
        // two CTAs of 256 threads per SM when two single staging buffers fit (C5: 2 x 80 KB), else one CTA of 512 threads
        // (16-lag groups keep 96 packed accumulator / ring registers: they run as one CTA of 384 threads with 168 registers)
        const bool twin = LB == 8 && 2 * (blk_bytes + 2 * (size_t)std::max<int64_t>(1, 256 / p2) * p * 4 + 512) <= 224 * 1024 &&
                          p2 <= 256 && !getenv("MMC_STATS_ONE_CTA");
        const int64_t tmax = twin ? 256 : (LB == 8 ? 512 : 384);
        const int64_t rmax = std::max<int64_t>(1, tmax / p2);           // replicas per (chain, pair) with one staged chain
        int H = (int)std::min<int64_t>((n_lags + LB - 1) / LB, rmax);
        if (H < 1) H = 1;
        int G = (int)(rmax / H);
        if (G < 1) G = 1;
        const int64_t steps = std::max<int64_t>(1, N - lag0);
        if (G > (int)((steps + LB - 1) / LB)) G = (int)((steps + LB - 1) / LB);   // a segment is at least one block of LB steps
        // small split-chain blocks (few parameters): prefer more staged chains per round over more time segments per
        // chain, so that a round moves >= 32 KB with its bulk copies (with K = 1 a 3 KB block per round is latency bound:
        // the all-lag pass of the C3 sample took 13 ms for 0.03 TMAC)
        const int64_t ktarget = std::min<int64_t>(16, ((int64_t)32 * 1024 + (int64_t)blk_bytes - 1) / (int64_t)blk_bytes);
        const int64_t gcap = std::max<int64_t>(1, tmax / (p2 * H * ktarget));
        if (G > gcap) G = (int)gcap;
        int K = (int)std::min<int64_t>(16, tmax / (p2 * H * G));
        if (K < 1) K = 1;
        const size_t budget = twin ? 112 * 1024 : 200 * 1024;
        while (K > 1 && (twin ? 1 : 2) * (size_t)K * blk_bytes + (size_t)H * G * K * p * 4 + 256 > budget) --K;
        if ((int64_t)K > 2 * c_local) K = (int)(2 * c_local);
        const size_t extra = (size_t)H * G * K * p * 4 + 256;
        const int nbuf = twin ? 1 : ((2 * K * blk_bytes + extra <= 220 * 1024) ? 2 : 1);
        const size_t smem = nbuf * K * blk_bytes + extra;
        const int threads = (int)((p2 * H * G * K + 31) / 32 * 32);
        MMC_REQUIRE(threads <= tmax, "stats: %d threads > %d", threads, (int)tmax);
        int64_t grid = (int64_t)sm_count() * (twin ? 2 : 1);
        if (grid > (2 * c_local + K - 1) / K) grid = (2 * c_local + K - 1) / K;
        void (*kern)(const float *, int64_t, int64_t, int, int64_t, int, int, int, int, double *);
        // kZero needs every replica on lag base 0: one lag group (H == 1) starting at lag 0
        if (twin) kern = stats_block2_kernel<8, true, true, 256, 2>;
        else if (LB == 8) kern = stats_block2_kernel<8, true, true, 512, 1>;
        else if (first) kern = H == 1 ? stats_block2_kernel<kLagBlock, true, true, 384, 1> : stats_block2_kernel<kLagBlock, true, false, 384, 1>;
        else kern = stats_block2_kernel<kLagBlock, false, false, 384, 1>;
        MMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)grid, threads, smem, stream>>>(sample, c_local, n, (int)p, lag0, H, G, K, nbuf, partial);
        MMC_CUDA(cudaGetLastError());
        *covered = std::min<int64_t>((int64_t)H * LB, N - lag0);
        return MMC_OK;
    }
    int H = (int)std::min<int64_t>((n_lags + kLagBlock - 1) / kLagBlock, 512 / p);
    if (H < 1) H = 1;
    int G = (int)(512 / ((int64_t)p * H));     // time segments: fill the CTA when few lag groups are requested
    if (G < 1) G = 1;
    if (G > (int)((N + kLagBlock - 1) / kLagBlock)) G = (int)((N + kLagBlock - 1) / kLagBlock);
    {   // small blocks: chains per round before time segments per chain (see the packed path above)
        const int64_t ktarget = std::min<int64_t>(16, ((int64_t)32 * 1024 + (int64_t)blk_bytes - 1) / (int64_t)blk_bytes);
        const int64_t gcap = std::max<int64_t>(1, 512 / (p * H * ktarget));
        if (G > gcap) G = (int)gcap;
    }
    // small p: stage K split chains per round so that the CTA still has ~512 threads of work
    int K = (int)std::min<int64_t>(16, 512 / ((int64_t)p * H * G));
    if (K < 1) K = 1;
    while (K > 1 && 2 * (size_t)K * blk_bytes + (size_t)H * G * K * p * 4 + 256 > 200 * 1024) --K;
    if ((int64_t)K > 2 * c_local) K = (int)(2 * c_local);
    const size_t extra = (size_t)H * G * K * p * 4 + 256;
    const int nbuf = (2 * K * blk_bytes + extra <= 220 * 1024) ? 2 : 1;
    const size_t smem = nbuf * K * blk_bytes + extra;
    int threads = (int)(((int64_t)H * G * K * p + 31) / 32 * 32);
    int64_t grid = sm_count();
    if (grid > (2 * c_local + K - 1) / K) grid = (2 * c_local + K - 1) / K;
    auto kern = lag0 == 0 ? stats_block_kernel<true> : stats_block_kernel<false>;
    MMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, threads, smem, stream>>>(sample, c_local, n, (int)p, lag0, H, G, K, nbuf, partial);
    MMC_CUDA(cudaGetLastError());
    *covered = std::min<int64_t>((int64_t)H * kLagBlock, N - lag0);
    return MMC_OK;
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

// ---------------------------------------------------------------- device-side finalisation
// Same arithmetic as mmc_stats_finalize (f64 moment algebra, f32 Geyer loop as in src/stats.rs:518-545), one thread per
// parameter, so that a round of the lag-window protocol costs one small D2H copy (rhat, ess, flag) instead of the whole
// partial.  ws[0] = number of LOCAL chains summed over the ranks, ws + 1 = the partial rows.
__global__ void stats_finalize_kernel(const double *__restrict__ ws, int64_t n, int p, int64_t lags_available,
                                      float *__restrict__ rhat_out, float *__restrict__ ess_out, int *__restrict__ need_more) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p) return;
    const double *partial = ws + 1;
    const int64_t N = n / 2;
    const double C = 2.0 * ws[0];
    if (lags_available > N) lags_available = N;
    const double sm = partial[q], sm2 = partial[p + q];
    const double mbar = sm / C;
    double ssd = sm2 - C * mbar * mbar;
    if (ssd < 0.0) ssd = 0.0;
    const double B = ssd * ((double)N / (C - 1.0));
    const double W = partial[2 * (int64_t)p + q] / C;
    const double var = (((double)N - 1.0) / (double)N) * W + B / (double)N;
    const float within = (float)W, varf = (float)var;
    rhat_out[q] = __fsqrt_rn(__fdiv_rn(within, varf));
    auto rho = [&](int64_t t) {
        const float avg = (float)(partial[(2 + t) * p + q] / C);
        return __fadd_rn(-__fdiv_rn(__fadd_rn(-avg, within), varf), 1.0f);
    };
    float mn = N >= 2 ? __fadd_rn(rho(0), rho(1)) : 0.0f;
    float o = 0.0f;
    bool terminated = false;
    int64_t t = 0;
    for (; t + 1 < lags_available; t += 2) {
        float pt = __fadd_rn(rho(t), rho(t + 1));
        if (pt <= 0.0f) { terminated = true; break; }
        if (pt > mn) pt = mn;
        mn = pt;
        o = __fadd_rn(o, pt);
    }
    if (!terminated && t + 1 < N) {
        // bit 0: more lags are needed.  bit 1 (first window only): the pair sums decay so slowly that the next, 64-lag window
        // cannot terminate Geyer's sum either (geometric extrapolation from the first and the last pair of this window), so
        // the protocol goes straight to all lags and the sample crosses HBM twice instead of three times (C5: 19.6 -> 13 ms)
        int f = 1;
        const int64_t pairs = t / 2;
        if (lags_available <= 8 && pairs >= 2) {
            const float p0 = __fadd_rn(rho(0), rho(1));
            if (p0 > 0.0f && mn > 0.0f) {
                const float r = powf(mn / p0, 1.0f / (float)(pairs - 1));   // decay per pair of lags
                if (r >= 0.98f || 2.0f * (logf(0.02f) / logf(r)) > 72.0f) f |= 2;
            }
        }
        atomicOr(need_more, f);
    }
    const float tau = __fadd_rn(-1.0f, __fmul_rn(2.0f, o));
    ess_out[q] = __fmul_rn(__fmul_rn(__fdiv_rn(1.0f, tau), (float)C), (float)N);
}

__global__ void stats_set_count_kernel(double *ws, double c_local, int *need_more) {
    ws[0] = c_local;
    *need_more = 0;
}

// ---------------------------------------------------------------- NCCL (resolved at run time)
// libminimcmc.so carries no link-time dependency on NCCL: a process that never shards its diagnostics does not need
// the library, and a process that already holds a copy (torch bundles its own libnccl.so.2) must keep using THAT copy.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    bool ok = false;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = nullptr;
        if (const char *path = getenv("MMC_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy this process already uses (e.g. torch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        bool all = true;
        auto sym = [&](const char *name) { void *f = dlsym(h, name); if (!f) all = false; return f; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.CommCount = reinterpret_cast<decltype(api.CommCount)>(sym("ncclCommCount"));
        api.CommUserRank = reinterpret_cast<decltype(api.CommUserRank)>(sym("ncclCommUserRank"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        api.ok = all;
    });
    return api.ok ? &api : nullptr;
}

int nccl_fail(NcclApi *api, ncclResult_t r, const char *what) {
    set_error("NCCL error %d (%s) in %s", (int)r, api && api->GetErrorString ? api->GetErrorString(r) : "?", what);
    return MMC_ERR_CUDA;
}

}  // namespace

struct mmc_comm {
    ncclComm_t comm = nullptr;
    bool owned = false;
    int nranks = 1, rank = 0;
};

namespace {

// grow-only per-process workspace of the diagnostics calls (cudaMalloc / cudaFree cost more than the kernels)
struct StatsWorkspace {
    std::mutex mutex;
    double *ws = nullptr;      // [1 + (2 + N) p] doubles: count slot, partial rows
    size_t ws_len = 0;
    float *out = nullptr;      // [2 p] floats (rhat, ess) + 1 int flag, device
    size_t out_len = 0;
    float *h_out = nullptr;    // pinned mirror
    int device = -1;
};
StatsWorkspace g_stats_ws;

// The lag-window protocol shared by the single-GPU and the sharded call (src/stats.rs:416-423 + SURVEY B3): local partial
// sums for a window of lags -> (sharded: ONE all-reduce of the new rows, the first one fused with the moment rows and the
// chain count) -> Geyer termination checked on the device -> another, twice as wide window only if some parameter has
// not terminated.
int split_rhat_ess_protocol(const float *sample_dev, int64_t c_local, int64_t n, int64_t p, mmc_comm *comm, cudaStream_t s,
                            float *rhat_host, float *ess_host) {
    const int64_t N = n / 2;
    const int64_t len = 1 + mmc_stats_partial_len(n, p);
    NcclApi *api = nullptr;
    if (comm && comm->nranks >= 1 && comm->comm) {
        api = nccl_api();
        MMC_REQUIRE(api, "NCCL is not available (libnccl.so.2 not found; set MMC_NCCL_LIB)");
    }
    StatsWorkspace &W = g_stats_ws;
    std::lock_guard<std::mutex> guard(W.mutex);
    int dev = 0;
    MMC_CUDA(cudaGetDevice(&dev));
    if (dev != W.device || (size_t)len > W.ws_len || (size_t)(2 * p + 1) > W.out_len) {
        if (W.ws) cudaFree(W.ws);
        if (W.out) cudaFree(W.out);
        if (W.h_out) cudaFreeHost(W.h_out);
        W.ws = nullptr; W.out = nullptr; W.h_out = nullptr; W.ws_len = W.out_len = 0; W.device = dev;
        MMC_CUDA(cudaMalloc((void **)&W.ws, sizeof(double) * (size_t)len));
        W.ws_len = (size_t)len;
        MMC_CUDA(cudaMalloc((void **)&W.out, sizeof(float) * (size_t)(2 * p + 1)));
        MMC_CUDA(cudaMallocHost((void **)&W.h_out, sizeof(float) * (size_t)(2 * p + 1)));
        W.out_len = (size_t)(2 * p + 1);
    }
    double *d_partial = W.ws + 1;
    int *d_flag = reinterpret_cast<int *>(W.out + 2 * p);
    // lag windows: 8 (well-mixed chains terminate there and the pass stays HBM bound), then 64 more, then everything
    int64_t have = 0, block = 8;
    while (have < N) {
        const int64_t want = std::min<int64_t>(block, N - have);
        if (have == 0) stats_set_count_kernel<<<1, 1, 0, s>>>(W.ws, (double)c_local, d_flag);
        else MMC_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), s));
        int rc = mmc_stats_partial_dev(sample_dev, c_local, n, p, have, want, d_partial, s);
        if (rc) return rc;
        if (api) {
            // round 0: count + moment rows + the first lag rows are one contiguous range; later rounds: the new lag rows
            double *buf = have == 0 ? W.ws : d_partial + (2 + have) * p;
            const size_t cnt = have == 0 ? (size_t)(1 + (2 + want) * p) : (size_t)(want * p);
            const ncclResult_t r = api->AllReduce(buf, buf, cnt, ncclFloat64, ncclSum, comm->comm, s);
            if (r != ncclSuccess) return nccl_fail(api, r, "ncclAllReduce (split-Rhat / ESS partial sums)");
        }
        have += want;
        stats_finalize_kernel<<<(unsigned)((p + 127) / 128), 128, 0, s>>>(W.ws, n, (int)p, have, W.out, W.out + p, d_flag);
        MMC_CUDA(cudaGetLastError());
        MMC_CUDA(cudaMemcpyAsync(W.h_out, W.out, sizeof(float) * (size_t)(2 * p + 1), cudaMemcpyDeviceToHost, s));
        MMC_CUDA(cudaStreamSynchronize(s));
        int flag;
        memcpy(&flag, W.h_out + 2 * p, sizeof(int));
        if (!flag) break;
        block = (have <= 8 && !(flag & 2)) ? 64 : N;  // at most three rounds; two when the first window predicts slow decay
    }
    if (rhat_host) memcpy(rhat_host, W.h_out, sizeof(float) * (size_t)p);
    if (ess_host) memcpy(ess_host, W.h_out + p, sizeof(float) * (size_t)p);
    return MMC_OK;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- communicator (NCCL over NVLink / NVSwitch)
int mmc_comm_unique_id(unsigned char *id128) {
    MMC_REQUIRE(id128, "mmc_comm_unique_id: null buffer");
    NcclApi *api = nccl_api();
    MMC_REQUIRE(api, "NCCL is not available (libnccl.so.2 not found; set MMC_NCCL_LIB)");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId uid;
    const ncclResult_t r = api->GetUniqueId(&uid);
    if (r != ncclSuccess) return nccl_fail(api, r, "ncclGetUniqueId");
    memcpy(id128, &uid, 128);
    return MMC_OK;
}

int mmc_comm_create(mmc_comm **out, const unsigned char *id128, int32_t nranks, int32_t rank) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(out && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "mmc_comm_create: bad arguments");
    NcclApi *api = nccl_api();
    MMC_REQUIRE(api, "NCCL is not available (libnccl.so.2 not found; set MMC_NCCL_LIB)");
    ncclUniqueId uid;
    memcpy(&uid, id128, 128);
    mmc_comm *c = new mmc_comm();
    const ncclResult_t r = api->CommInitRank(&c->comm, nranks, uid, rank);   // binds the calling thread's current device
    if (r != ncclSuccess) { delete c; return nccl_fail(api, r, "ncclCommInitRank"); }
    c->owned = true;
    c->nranks = nranks;
    c->rank = rank;
    *out = c;
    return MMC_OK;
}

int mmc_comm_wrap(mmc_comm **out, void *nccl_comm) {
    MMC_REQUIRE(out && nccl_comm, "mmc_comm_wrap: bad arguments");
    NcclApi *api = nccl_api();
    MMC_REQUIRE(api, "NCCL is not available (libnccl.so.2 not found; set MMC_NCCL_LIB)");
    mmc_comm *c = new mmc_comm();
    c->comm = static_cast<ncclComm_t>(nccl_comm);
    ncclResult_t r = api->CommCount(c->comm, &c->nranks);
    if (r == ncclSuccess) r = api->CommUserRank(c->comm, &c->rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(api, r, "ncclCommCount"); }
    *out = c;
    return MMC_OK;
}

int mmc_comm_info(mmc_comm *c, int32_t *nranks, int32_t *rank, int32_t *nccl_version) {
    MMC_REQUIRE(c, "null communicator");
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    if (nccl_version) {
        int v = 0;
        NcclApi *api = nccl_api();
        if (api) api->GetVersion(&v);
        *nccl_version = v;
    }
    return MMC_OK;
}

void mmc_comm_destroy(mmc_comm *c) {
    if (!c) return;
    NcclApi *api = nccl_api();
    if (c->owned && c->comm && api) api->CommDestroy(c->comm);
    delete c;
}

int mmc_split_rhat_ess_sharded(const float *sample_dev, int64_t c_local, int64_t n, int64_t p, mmc_comm *comm, void *stream,
                               float *rhat_host, float *ess_host) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(sample_dev && c_local > 0 && n >= 2 && p > 0 && comm, "mmc_split_rhat_ess_sharded: bad arguments");
    return split_rhat_ess_protocol(sample_dev, c_local, n, p, comm, (cudaStream_t)stream, rhat_host, ess_host);
}

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
    {   // zero the rows this call produces (the staged kernel rounds its last group up to 16 lags)
        const int64_t hi = std::min<int64_t>(N, lag0 + (n_lags + kLagBlock - 1) / kLagBlock * kLagBlock);
        MMC_CUDA(cudaMemsetAsync(partial_dev + (2 + lag0) * p, 0, sizeof(double) * (hi - lag0) * p, s));
    }
    for (int64_t l = 0; l < n_lags;) {
        int64_t covered = 0;
        rc = launch_block_pass(sample_dev, c_local, n, p, lag0 + l, n_lags - l, partial_dev, s, &covered);
        if (rc) return rc;
        if (covered > 0) {   // shared-memory staged variant handled `covered` lags (it may compute a few beyond
            l += covered;    // n_lags inside its last group of 16; those rows were zeroed below)
            continue;
        }
        const int nl = (int)std::min<int64_t>(kLagBlock, n_lags - l);
        rc = launch_pass(sample_dev, c_local, n, p, lag0 + l, nl, partial_dev, s);
        if (rc) return rc;
        l += kLagBlock;
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
    return split_rhat_ess_protocol(sample_dev, c, n, p, nullptr, (cudaStream_t)stream, rhat_host, ess_host);
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
