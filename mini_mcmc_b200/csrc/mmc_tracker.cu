// Device-resident progress trackers: the statistics behind run_progress.
//   flavor MMC_TRACK_MULTI     = MultiChainTracker        src/stats.rs:189-307 (HMC::run_progress, src/hmc.rs:242-281)
//   flavor MMC_TRACK_PER_CHAIN = one ChainTracker / chain src/stats.rs:26-141 + collect_rhat :150-178
//                                (ChainRunner::run_progress, src/core.rs:90-136,229-324)
// The reference copies every step's state to the host and folds it into f32 running moments there.  Here the
// draws never leave HBM: the samplers write [chains, n, dim] blocks of draws, and one streaming pass per block
// replays the same f32 recurrences (mean_n = (mean_{n-1} (n-1) + x) / n, same for x^2; never contracted: this
// file is compiled with -fmad=false) with one lane per chain (dim <= 8, rows staged through shared memory so the
// global reads stay coalesced) or one warp per chain (lanes over the parameters).  The cross-chain sums behind
// Rhat are f64 partials that can be all-reduced over ranks before the (f32) finalisation.
#include <algorithm>
#include <vector>

#include "mmc_common.cuh"
#include "mmc_stepdiv.cuh"

struct mmc_tracker {
    int64_t chains = 0;
    int32_t dim = 0, flavor = 0;
    uint64_t n = 0;
    float *d_mean = nullptr, *d_msq = nullptr, *d_last = nullptr, *d_p_chain = nullptr, *d_p_global = nullptr;
    uint8_t *d_flags = nullptr;
    double *d_partial = nullptr;
    cudaStream_t stream = nullptr;  // used by the synchronous entry points until a *_dev call names the caller's stream
    cudaStream_t last = nullptr;    // stream of the most recent mmc_tracker_steps_dev (summary/get order themselves after it)
    bool have_last = false;
};

namespace mmc {
namespace {

constexpr int kWindow = 16384;  // accept-EMA window: 0.99^16384 underflows f32, older flags cannot matter
constexpr float kAlpha = 0.01f; // src/stats.rs:13

struct TrackParams {
    const void *sample;  // [chains, n_total, dim]
    int64_t n_total, t0, k, chains;
    int32_t dim, flavor;
    float *mean, *msq, *last, *p_chain;
    uint8_t *flags;      // last kWindow accept flags of this block in (step, chain) order (MULTI)
    int64_t flag_base;   // linear index (t_rel * chains + c) of flags[0]
    uint64_t n_before;   // steps folded before this block
};

template <typename InT> __device__ __forceinline__ float to_f32(InT v) { return (float)v; }

// the divider is StepDiv (one reciprocal per step shared by the NC chunks of a lane: wide kernel) or IeeeDiv (lane-per-chain
// kernel, where a step only has 2 dim divisions and the reciprocal would not pay)
struct IeeeDiv {
    float n;
    __device__ __forceinline__ explicit IeeeDiv(float n_) : n(n_) {}
    __device__ __forceinline__ float operator()(float a) const { return __fdiv_rn(a, n); }
};
template <class Div>
__device__ __forceinline__ void fold_moments(float &mean, float &msq, float x, const Div &dv, float nm1, bool first) {
    mean = dv(__fadd_rn(__fmul_rn(mean, nm1), x));
    const float xx = __fmul_rn(x, x);
    msq = first ? xx : dv(__fadd_rn(__fmul_rn(msq, nm1), xx));
}

__device__ __forceinline__ float fold_accept(float p, bool accepted) {
    return __fadd_rn(__fmul_rn(__fsub_rn(1.0f, kAlpha), p), __fmul_rn(kAlpha, accepted ? 1.0f : 0.0f));
}

// ---- dim <= 8: lane per chain, 32 chains per warp, rows staged through shared memory
constexpr int kSmallWarps = 4;
template <typename InT> struct SmallRow { static constexpr int kElems = 256 / (int)sizeof(InT); };  // 256 B staged per chain per round

template <typename InT>
__global__ void __launch_bounds__(kSmallWarps * 32) tracker_small_kernel(const TrackParams p) {
    constexpr int kSmallRow = SmallRow<InT>::kElems;
    __shared__ InT tile[kSmallWarps][32][kSmallRow + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c_base = ((int64_t)blockIdx.x * kSmallWarps + warp) * 32;
    if (c_base >= p.chains) return;
    const int64_t c = c_base + lane;
    const bool live = c < p.chains;
    const int dim = p.dim;
    const int steps_per_round = kSmallRow / dim;
    const InT *sample = static_cast<const InT *>(p.sample);

    float mean[8], msq[8], last[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        const bool ok = live && d < dim;
        mean[d] = ok ? p.mean[c * dim + d] : 0.0f;
        msq[d] = ok ? p.msq[c * dim + d] : 0.0f;
        last[d] = ok ? p.last[c * dim + d] : 0.0f;
    }
    float pc = (live && p.flavor == MMC_TRACK_PER_CHAIN) ? p.p_chain[c] : 0.0f;

    for (int64_t tr = 0; tr < p.k; tr += steps_per_round) {
        const int ns = (int)min((int64_t)steps_per_round, p.k - tr);
        const int row_len = ns * dim;
        __syncwarp();
        for (int r = 0; r < 32; ++r) {
            const int64_t cr = c_base + r;
            if (cr >= p.chains) break;
            const InT *src = sample + (cr * p.n_total + p.t0 + tr) * dim;
            for (int j = lane; j < row_len; j += 32) tile[warp][r][j] = __ldcs(src + j);
        }
        __syncwarp();
        if (live) {
            for (int s = 0; s < ns; ++s) {
                const uint64_t step = p.n_before + (uint64_t)(tr + s) + 1;
                const float n = (float)step;
                const IeeeDiv dv(n);
                const float nm1 = __fsub_rn(n, 1.0f);
                bool changed = false, changed0 = false;
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    if (d < dim) {
                        const float x = to_f32(tile[warp][lane][s * dim + d]);
                        fold_moments(mean[d], msq[d], x, dv, nm1, step == 1);
                        const bool ne = x != last[d];
                        changed |= ne;
                        if (d == 0) changed0 = ne;
                        last[d] = x;
                    }
                }
                if (p.flavor == MMC_TRACK_PER_CHAIN) {
                    if (pc < 0.0f) pc = changed0 ? 1.0f : 0.0f;  // src/stats.rs:107-114
                    pc = fold_accept(pc, changed);
                } else {
                    const int64_t i = (tr + s) * p.chains + c - p.flag_base;
                    if (i >= 0) p.flags[i] = changed ? 1 : 0;
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            if (d < dim) {
                p.mean[c * dim + d] = mean[d];
                p.msq[c * dim + d] = msq[d];
                p.last[c * dim + d] = last[d];
            }
        }
        if (p.flavor == MMC_TRACK_PER_CHAIN) p.p_chain[c] = pc;
    }
}

// ---- dim > 8: warp per chain, lanes over the parameters, 32 steps per round.  NC chunks of 32 parameters advance together,
// so a row (dim contiguous values) is read once, by adjacent loads of one step: sectors that straddle two chunks or two rows
// are L1 hits instead of a second trip to L2 / HBM (400 B rows: 1.41x read amplification before), and the NC recurrences of
// a lane are independent instruction streams.
constexpr int kWideWarps = 8;

template <typename InT, int NC>
__global__ void __launch_bounds__(kWideWarps * 32) tracker_wide_kernel(const TrackParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * kWideWarps + (threadIdx.x >> 5);
    if (c >= p.chains) return;
    const int dim = p.dim;
    const InT *row = static_cast<const InT *>(p.sample) + (c * p.n_total + p.t0) * dim;
    float pc = p.flavor == MMC_TRACK_PER_CHAIN ? p.p_chain[c] : 0.0f;

    for (int64_t tr = 0; tr < p.k; tr += 32) {
        const int ns = (int)min((int64_t)32, p.k - tr);
        uint32_t changed_mask = 0, changed0_mask = 0;
        for (int d0 = 0; d0 < dim; d0 += 32 * NC) {
            float mean[NC], msq[NC], last[NC];
            bool ok[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const int d = d0 + 32 * j + lane;
                ok[j] = d < dim;
                mean[j] = ok[j] ? p.mean[c * dim + d] : 0.0f;
                msq[j] = ok[j] ? p.msq[c * dim + d] : 0.0f;
                last[j] = ok[j] ? p.last[c * dim + d] : 0.0f;
            }
            const InT *src = row + tr * dim + d0 + lane;
            const uint64_t step0 = p.n_before + (uint64_t)tr + 1;   // step count of s = 0
            const bool small = step0 + 32 < (1ull << 32);            // (always, in practice): 32-bit int -> float conversions
            const uint32_t step0_lo = (uint32_t)step0;
            uint32_t ne_bits = 0, ne0_bits = 0;                     // per lane: bit s = this lane's element changed at step s
#pragma unroll 4
            for (int s = 0; s < ns; ++s) {
                const float n = small ? (float)(step0_lo + (uint32_t)s) : (float)(step0 + (uint64_t)s);
                const StepDiv dv(n);
                const float nm1 = __fsub_rn(n, 1.0f);
                const bool first = step0 + (uint64_t)s == 1;
                float x[NC];
#pragma unroll
                for (int j = 0; j < NC; ++j) x[j] = ok[j] ? to_f32(__ldcs(src + 32 * j)) : 0.0f;
                src += dim;
                // the 2 NC numerators of this step (the recurrences of fold_moments); one range test and ONE branch for all
                float am[NC], aq[NC];
                bool ne = false, all_in = dv.fast;
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    am[j] = __fadd_rn(__fmul_rn(mean[j], nm1), x[j]);
                    aq[j] = __fadd_rn(__fmul_rn(msq[j], nm1), __fmul_rn(x[j], x[j]));
                    if (ok[j]) all_in = all_in && StepDiv::in_range(am[j]) && (first || StepDiv::in_range(aq[j]));
                    const bool d = ok[j] && x[j] != last[j];
                    ne |= d;
                    if (j == 0) ne0_bits |= (uint32_t)d << s;
                    last[j] = x[j];
                }
                if (all_in) {
#pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        mean[j] = dv.quotient_in_range(am[j]);
                        msq[j] = first ? __fmul_rn(x[j], x[j]) : dv.quotient_in_range(aq[j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        mean[j] = __fdiv_rn(am[j], n);
                        msq[j] = first ? __fmul_rn(x[j], x[j]) : __fdiv_rn(aq[j], n);
                    }
                }
                ne_bits |= (uint32_t)ne << s;
            }
            changed_mask |= __reduce_or_sync(0xffffffffu, ne_bits);
            if (d0 == 0) changed0_mask |= __shfl_sync(0xffffffffu, ne0_bits, 0);
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                if (ok[j]) {
                    const int d = d0 + 32 * j + lane;
                    p.mean[c * dim + d] = mean[j];
                    p.msq[c * dim + d] = msq[j];
                    p.last[c * dim + d] = last[j];
                }
            }
        }
        if (p.flavor == MMC_TRACK_PER_CHAIN) {
            for (int s = 0; s < ns; ++s) {
                if (pc < 0.0f) pc = ((changed0_mask >> s) & 1u) ? 1.0f : 0.0f;
                pc = fold_accept(pc, (changed_mask >> s) & 1u);
            }
        } else if (lane < ns) {
            const int64_t i = (tr + lane) * p.chains + c - p.flag_base;
            if (i >= 0) p.flags[i] = (changed_mask >> lane) & 1u;
        }
    }
    if (p.flavor == MMC_TRACK_PER_CHAIN && lane == 0) p.p_chain[c] = pc;
}

// MultiChainTracker::step folds the accept flags of all chains, in chain order, every step (src/stats.rs:248-256):
// one sequential EMA over the (step, chain)-ordered flags.  Only the last kWindow of them can influence an f32.
__global__ void tracker_ema_kernel(const uint8_t *flags, int64_t count, int restart, float *p_global) {
    float pa = restart ? 0.0f : *p_global;
    for (int64_t i = 0; i < count; ++i) pa = fold_accept(pa, flags[i] != 0);
    *p_global = pa;
}

// f64 partial sums over the local chains, per parameter d: [sum mean, sum mean^2, sum sm2] + [sum p_chain]
__global__ void __launch_bounds__(256) tracker_partial_kernel(const float *mean, const float *msq, const float *p_chain, int64_t chains,
                                                              int32_t dim, int64_t threads_total, uint64_t n_steps, double *partial) {
    extern __shared__ double acc[];  // [3 * dim] when dim <= 256
    const bool use_smem = dim <= 256;
    if (use_smem) {
        for (int i = threadIdx.x; i < 3 * dim; i += blockDim.x) acc[i] = 0.0;
        __syncthreads();
    }
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < threads_total) {
        const int d = (int)(tid % dim);
        const float n = (float)n_steps, nm1 = __fsub_rn(n, 1.0f);
        double s1 = 0.0, s2 = 0.0, s3 = 0.0, sp = 0.0;
        const int64_t total = chains * dim;
        for (int64_t e = tid; e < total; e += threads_total) {
            const float m = mean[e], q = msq[e];
            // sm2 = (mean_sq - mean^2) * n / (n - 1), src/stats.rs:138,297
            const float sm2 = __fdiv_rn(__fmul_rn(__fsub_rn(q, __fmul_rn(m, m)), n), nm1);
            s1 += (double)m;
            s2 += (double)m * (double)m;
            s3 += (double)sm2;
            if (d == 0 && p_chain) sp += (double)p_chain[e / dim];
        }
        double *dst = use_smem ? acc : partial;
        atomicAdd(dst + d, s1);
        atomicAdd(dst + dim + d, s2);
        atomicAdd(dst + 2 * dim + d, s3);
        if (d == 0 && p_chain) atomicAdd(partial + 3 * dim, sp);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * dim; i += blockDim.x) atomicAdd(partial + i, acc[i]);
    }
}

// slot 3*dim = chain-weighted accept estimate (MULTI: p_accept * chains, so a cross-rank sum / chains_total stays an average and a
// single rank gets the reference's value back exactly), slot 3*dim + 1 = chains
__global__ void tracker_pack_kernel(const float *p_global, int64_t chains, int32_t dim, int multi, double *partial) {
    if (multi) partial[3 * dim] = (double)*p_global * (double)chains;
    partial[3 * dim + 1] = (double)chains;
}

}  // namespace
}  // namespace mmc

using namespace mmc;


extern "C" {

void mmc_tracker_destroy(mmc_tracker *t) {
    if (!t) return;
    cudaFree(t->d_mean); cudaFree(t->d_msq); cudaFree(t->d_last); cudaFree(t->d_p_chain); cudaFree(t->d_p_global);
    cudaFree(t->d_flags); cudaFree(t->d_partial);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

int mmc_tracker_create(mmc_tracker **out, int64_t chains, int32_t dim, int32_t flavor) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(out && chains > 0 && dim > 0, "mmc_tracker_create: bad arguments");
    MMC_REQUIRE(flavor == MMC_TRACK_MULTI || flavor == MMC_TRACK_PER_CHAIN, "mmc_tracker_create: unknown flavor %d", flavor);
    mmc_tracker *t = new mmc_tracker();
    t->chains = chains; t->dim = dim; t->flavor = flavor;
    const size_t nb = (size_t)chains * dim * sizeof(float);
    auto fail = [&](cudaError_t e, const char *what) { int code = cuda_fail(e, what, __FILE__, __LINE__); mmc_tracker_destroy(t); return code; };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if ((e = cudaMalloc(&t->d_mean, nb)) != cudaSuccess) return fail(e, "cudaMalloc(mean)");
    if ((e = cudaMalloc(&t->d_msq, nb)) != cudaSuccess) return fail(e, "cudaMalloc(mean_sq)");
    if ((e = cudaMalloc(&t->d_last, nb)) != cudaSuccess) return fail(e, "cudaMalloc(last_state)");
    if ((e = cudaMalloc(&t->d_p_chain, chains * sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(p_accept)");
    if ((e = cudaMalloc(&t->d_p_global, sizeof(float))) != cudaSuccess) return fail(e, "cudaMalloc(p_accept)");
    if ((e = cudaMalloc(&t->d_flags, kWindow)) != cudaSuccess) return fail(e, "cudaMalloc(flags)");
    if ((e = cudaMalloc(&t->d_partial, (3 * (size_t)dim + 2) * sizeof(double))) != cudaSuccess) return fail(e, "cudaMalloc(partial)");
    // MultiChainTracker::new: zeros, p_accept 0 (src/stats.rs:208-220); ChainTracker::new: p_accept -1 (src/stats.rs:59-80)
    cudaMemsetAsync(t->d_mean, 0, nb, t->stream);
    cudaMemsetAsync(t->d_msq, 0, nb, t->stream);
    cudaMemsetAsync(t->d_last, 0, nb, t->stream);
    cudaMemsetAsync(t->d_p_global, 0, sizeof(float), t->stream);
    std::vector<float> minus_one((size_t)chains, flavor == MMC_TRACK_PER_CHAIN ? -1.0f : 0.0f);
    e = cudaMemcpyAsync(t->d_p_chain, minus_one.data(), chains * sizeof(float), cudaMemcpyHostToDevice, t->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
    if (e != cudaSuccess) return fail(e, "tracker init");
    *out = t;
    return MMC_OK;
}

}  // extern "C"

namespace mmc {
namespace {

template <typename InT>
__global__ void tracker_cast_kernel(const InT *src, float *dst, int64_t len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) dst[i] = (float)src[i];
}

template <typename InT>
int launch_steps(mmc_tracker *t, const TrackParams &p, cudaStream_t s) {
    if (t->dim <= 8) {
        const int64_t warps = (t->chains + 31) / 32;
        tracker_small_kernel<InT><<<(unsigned)((warps + kSmallWarps - 1) / kSmallWarps), kSmallWarps * 32, 0, s>>>(p);
    } else {
        const unsigned grid = (unsigned)((t->chains + kWideWarps - 1) / kWideWarps);
        if (t->dim <= 32) tracker_wide_kernel<InT, 1><<<grid, kWideWarps * 32, 0, s>>>(p);
        else if (t->dim <= 64) tracker_wide_kernel<InT, 2><<<grid, kWideWarps * 32, 0, s>>>(p);
        else if (t->dim <= 96) tracker_wide_kernel<InT, 3><<<grid, kWideWarps * 32, 0, s>>>(p);
        else tracker_wide_kernel<InT, 4><<<grid, kWideWarps * 32, 0, s>>>(p);
    }
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

}  // namespace
}  // namespace mmc

extern "C" {

int mmc_tracker_set_initial_dev(mmc_tracker *t, const void *state_dev, int32_t dtype, void *stream) {
    MMC_REQUIRE(t && state_dev, "mmc_tracker_set_initial_dev: bad arguments");
    MMC_REQUIRE(t->n == 0, "mmc_tracker_set_initial_dev: the tracker has already folded %llu steps", (unsigned long long)t->n);
    const int64_t len = t->chains * t->dim;
    const unsigned grid = (unsigned)((len + 255) / 256);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == MMC_F32) tracker_cast_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float *>(state_dev), t->d_last, len);
    else if (dtype == MMC_F64) tracker_cast_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double *>(state_dev), t->d_last, len);
    else if (dtype == MMC_U64) tracker_cast_kernel<unsigned long long><<<grid, 256, 0, s>>>(static_cast<const unsigned long long *>(state_dev), t->d_last, len);
    else MMC_REQUIRE(false, "mmc_tracker_set_initial_dev: unknown dtype %d", dtype);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

int mmc_tracker_steps_dev(mmc_tracker *t, const void *sample_dev, int32_t dtype, int64_t n_total, int64_t t0,
                          int64_t n_steps, void *stream) {
    MMC_REQUIRE(t && sample_dev && n_total > 0 && t0 >= 0 && n_steps >= 0 && t0 + n_steps <= n_total,
                "mmc_tracker_steps_dev: bad arguments");
    if (n_steps == 0) return MMC_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TrackParams p{};
    p.sample = sample_dev;
    p.n_total = n_total; p.t0 = t0; p.k = n_steps; p.chains = t->chains;
    p.dim = t->dim; p.flavor = t->flavor;
    p.mean = t->d_mean; p.msq = t->d_msq; p.last = t->d_last; p.p_chain = t->d_p_chain;
    p.flags = t->d_flags;
    const int64_t count = n_steps * t->chains;  // chains * steps < 2^63 for any tensor that fits HBM
    p.flag_base = std::max<int64_t>(0, count - kWindow);
    p.n_before = t->n;
    int rc;
    if (dtype == MMC_F32) rc = launch_steps<float>(t, p, s);
    else if (dtype == MMC_F64) rc = launch_steps<double>(t, p, s);
    else if (dtype == MMC_U64) rc = launch_steps<unsigned long long>(t, p, s);
    else { set_error("mmc_tracker_steps_dev: unknown dtype %d", dtype); return MMC_ERR_INVALID; }
    if (rc) return rc;
    if (t->flavor == MMC_TRACK_MULTI) {
        tracker_ema_kernel<<<1, 1, 0, s>>>(t->d_flags, std::min<int64_t>(count, kWindow), count >= kWindow ? 1 : 0, t->d_p_global);
        MMC_CUDA(cudaGetLastError());
    }
    t->n += (uint64_t)n_steps;
    t->last = s;
    t->have_last = true;
    return MMC_OK;
}

int64_t mmc_tracker_partial_len(int32_t dim) { return 3 * (int64_t)dim + 2; }

int mmc_tracker_partial_dev(mmc_tracker *t, double *partial_dev, void *stream) {
    MMC_REQUIRE(t && partial_dev, "mmc_tracker_partial_dev: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int dim = t->dim;
    MMC_CUDA(cudaMemsetAsync(partial_dev, 0, (size_t)mmc_tracker_partial_len(dim) * sizeof(double), s));
    const int64_t total = t->chains * dim;
    const int64_t want = std::min<int64_t>(total, (int64_t)sm_count() * 2048);
    const int64_t threads_total = ((want + dim - 1) / dim) * dim;  // multiple of dim: a thread keeps one parameter
    const unsigned grid = (unsigned)((threads_total + 255) / 256);
    const size_t smem = dim <= 256 ? 3 * (size_t)dim * sizeof(double) : 0;
    tracker_partial_kernel<<<grid, 256, smem, s>>>(t->d_mean, t->d_msq, t->flavor == MMC_TRACK_PER_CHAIN ? t->d_p_chain : nullptr,
                                                   t->chains, dim, threads_total, t->n, partial_dev);
    MMC_CUDA(cudaGetLastError());
    tracker_pack_kernel<<<1, 1, 0, s>>>(t->d_p_global, t->chains, dim, t->flavor == MMC_TRACK_MULTI ? 1 : 0, partial_dev);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

int mmc_tracker_finalize(const double *partial_host, int64_t chains_total, int32_t dim, uint64_t n_steps, int32_t flavor,
                         float *rhat_host, float *max_rhat) {
    MMC_REQUIRE(partial_host && chains_total > 0 && dim > 0, "mmc_tracker_finalize: bad arguments");
    const float n = (float)n_steps, c = (float)chains_total;
    float mx = -INFINITY;
    for (int d = 0; d < dim; ++d) {
        const double s1 = partial_host[d], s2 = partial_host[dim + d], s3 = partial_host[2 * dim + d];
        const double mbar = s1 / (double)chains_total;
        const double ssd = std::max(0.0, s2 - (double)chains_total * mbar * mbar);  // sum_c (mean_c - mean_bar)^2
        const float within = (float)(s3 / (double)chains_total);
        float var;
        if (flavor == MMC_TRACK_MULTI) {
            // src/stats.rs:287-303
            const float between = (float)ssd * (n / (c - 1.0f));
            var = within * ((n - 1.0f) / n) + between * (1.0f / n);
        } else {
            // withinvar_from_cs, src/stats.rs:155-178: between divides by chains * params - 1
            const float between = (float)ssd / (float)(chains_total * dim - 1);
            var = between + within * ((n - 1.0f) / n);
        }
        const float r = sqrtf(var / within);
        if (rhat_host) rhat_host[d] = r;
        if (r > mx) mx = r;  // max_skipnan
    }
    if (max_rhat) *max_rhat = mx;
    return MMC_OK;
}

int mmc_tracker_summary(mmc_tracker *t, float *rhat_host, float *max_rhat, float *p_accept, uint64_t *n_steps) {
    MMC_REQUIRE(t, "mmc_tracker_summary: null tracker");
    cudaStream_t s = t->have_last ? t->last : t->stream;
    int rc = mmc_tracker_partial_dev(t, t->d_partial, s);
    if (rc) return rc;
    std::vector<double> h((size_t)mmc_tracker_partial_len(t->dim));
    MMC_CUDA(cudaMemcpyAsync(h.data(), t->d_partial, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    MMC_CUDA(cudaStreamSynchronize(s));
    if (p_accept) *p_accept = (float)(h[3 * (size_t)t->dim] / (double)t->chains);
    if (n_steps) *n_steps = t->n;
    return mmc_tracker_finalize(h.data(), t->chains, t->dim, t->n, t->flavor, rhat_host, max_rhat);
}

int mmc_tracker_get(mmc_tracker *t, float *mean_host, float *mean_sq_host, float *p_accept_chain_host) {
    MMC_REQUIRE(t, "mmc_tracker_get: null tracker");
    const size_t nb = (size_t)t->chains * t->dim * sizeof(float);
    cudaStream_t s = t->have_last ? t->last : t->stream;
    if (mean_host) MMC_CUDA(cudaMemcpyAsync(mean_host, t->d_mean, nb, cudaMemcpyDeviceToHost, s));
    if (mean_sq_host) MMC_CUDA(cudaMemcpyAsync(mean_sq_host, t->d_msq, nb, cudaMemcpyDeviceToHost, s));
    if (p_accept_chain_host)
        MMC_CUDA(cudaMemcpyAsync(p_accept_chain_host, t->flavor == MMC_TRACK_PER_CHAIN ? t->d_p_chain : t->d_p_global,
                                 (t->flavor == MMC_TRACK_PER_CHAIN ? t->chains : 1) * sizeof(float), cudaMemcpyDeviceToHost, s));
    MMC_CUDA(cudaStreamSynchronize(s));
    return MMC_OK;
}

}  // extern "C"
