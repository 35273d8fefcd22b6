// run_progress at the C-ABI level: ChainRunner::run_progress (src/core.rs:208-360), HMC::run_progress
// (src/hmc.rs:222-294), NUTS::run_progress (src/nuts.rs:194-338).  The reference copies every step's state to the
// host and updates its trackers there; here the sampler runs in blocks of steps that write straight into windows of
// the final [chains, n_collect, dim] tensor in HBM, the device tracker (mmc_tracker.cu) folds each block, and the
// caller's callback receives p(accept) / max(rhat) once per block.  RunStats comes from the device split-Rhat / ESS.
#pragma once
#include <limits>

#include <algorithm>
#include <vector>

#include "mmc_common.cuh"

namespace mmc {

struct ProgressSpec {
    int64_t chains;
    int32_t dim;
    int32_t dtype;            // mmc_dtype of the draws
    int32_t flavor;           // MMC_TRACK_MULTI (HMC) or MMC_TRACK_PER_CHAIN (ChainRunner / NUTS)
    bool track_burn_in;       // ChainRunner / NUTS trackers see every step; HMC's tracker starts after burn-in
    const void *state_dev;    // current [chains, dim] state of the sampler (same dtype)
};

template <typename InT>
__global__ void progress_cast_kernel(const InT *__restrict__ src, float *__restrict__ dst, int64_t len) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) dst[i] = (float)src[i];
}

// run_block(k, dst_dev, pitch_steps, first) runs k steps of the sampler writing draws to dst_dev (rows of pitch_steps
// draws per chain); discard(k) runs k untracked steps.
template <class RunBlock, class Discard>
int run_progress_blocks(const ProgressSpec &sp, int64_t n_collect, int64_t n_discard, void *out_host, int64_t block,
                        mmc_progress_fn cb, void *user, mmc_run_stats *stats, cudaStream_t stream, RunBlock run_block, Discard discard) {
    if (stats) {   // RunStats is always returned; undefined entries (fewer than two collected draws) are NaN like the reference's
        const float nan = std::numeric_limits<float>::quiet_NaN();
        stats->ess = mmc_basic_stats{nan, nan, nan, nan, nan};
        stats->rhat = mmc_basic_stats{nan, nan, nan, nan, nan};
    }
    const size_t esize = sp.dtype == MMC_F32 ? 4 : 8;
    const int64_t total = n_collect + n_discard;
    const int64_t tracked_total = sp.track_burn_in ? total : n_collect;
    if (block <= 0) block = std::max<int64_t>(32, (tracked_total + 15) / 16);
    block = (block + 31) / 32 * 32;   // multiples of 32 draws keep the vectorised stores of the samplers aligned
    char *d_sample = nullptr, *d_scratch = nullptr;
    float *d_f32 = nullptr;
    mmc_tracker *tr = nullptr;
    int rc = MMC_OK;
    auto cleanup = [&]() {
        if (tr) mmc_tracker_destroy(tr);
        cudaFree(d_sample);
        cudaFree(d_scratch);
        cudaFree(d_f32);
    };
    auto fail_cuda = [&](cudaError_t e, const char *what) { int code = cuda_fail(e, what, __FILE__, __LINE__); cleanup(); return code; };
    const size_t sample_bytes = (size_t)sp.chains * n_collect * sp.dim * esize;
    cudaError_t e = cudaMalloc((void **)&d_sample, sample_bytes ? sample_bytes : 8);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(run_progress sample)");
    if ((rc = mmc_tracker_create(&tr, sp.chains, sp.dim, sp.flavor))) { cleanup(); return rc; }
    auto report = [&](int64_t done) -> int {
        if (!cb) return MMC_OK;
        float mx = 0.f, pa = 0.f;
        uint64_t n = 0;
        int r = mmc_tracker_summary(tr, nullptr, &mx, &pa, &n);
        if (r) return r;
        cb(done, tracked_total, pa, mx, user);
        return MMC_OK;
    };
    if (sp.track_burn_in) {
        // ChainTracker::new(initial state), then every step of burn-in through a scratch block (src/core.rs:104-121)
        if ((rc = mmc_tracker_set_initial_dev(tr, sp.state_dev, sp.dtype, stream))) { cleanup(); return rc; }
        if (n_discard > 0) {
            const int64_t blk = std::min(block, n_discard);
            e = cudaMalloc((void **)&d_scratch, (size_t)sp.chains * blk * sp.dim * esize);
            if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(run_progress scratch)");
            bool first = true;
            for (int64_t t0 = 0; t0 < n_discard; t0 += blk) {
                const int64_t k = std::min(blk, n_discard - t0);
                if ((rc = run_block(k, d_scratch, blk, first))) { cleanup(); return rc; }
                first = false;
                if ((rc = mmc_tracker_steps_dev(tr, d_scratch, sp.dtype, blk, 0, k, stream))) { cleanup(); return rc; }
                if ((rc = report(t0 + k))) { cleanup(); return rc; }
            }
        }
    } else {
        // HMC: burn-in is not tracked; the tracker first folds the post-burn-in positions (src/hmc.rs:228-249)
        if (n_discard > 0 && (rc = discard(n_discard))) { cleanup(); return rc; }
        if ((rc = mmc_tracker_steps_dev(tr, sp.state_dev, sp.dtype, 1, 0, 1, stream))) { cleanup(); return rc; }
    }
    bool first = !(sp.track_burn_in && n_discard > 0);
    for (int64_t t0 = 0; t0 < n_collect; t0 += block) {
        const int64_t k = std::min(block, n_collect - t0);
        if ((rc = run_block(k, d_sample + (size_t)t0 * sp.dim * esize, n_collect, first))) { cleanup(); return rc; }
        first = false;
        if ((rc = mmc_tracker_steps_dev(tr, d_sample, sp.dtype, n_collect, t0, k, stream))) { cleanup(); return rc; }
        if ((rc = report((sp.track_burn_in ? n_discard : 0) + t0 + k))) { cleanup(); return rc; }
    }
    if (out_host && sample_bytes) {
        e = cudaMemcpyAsync(out_host, d_sample, sample_bytes, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return fail_cuda(e, "run_progress D2H");
    }
    if (stats && n_collect >= 2) {
        // RunStats::from(sample): split-Rhat / ESS of the f32 view (src/stats.rs:352-371)
        const float *f32 = reinterpret_cast<const float *>(d_sample);
        const int64_t len = sp.chains * n_collect * sp.dim;
        if (sp.dtype != MMC_F32) {
            e = cudaMalloc((void **)&d_f32, (size_t)len * sizeof(float));
            if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(run_progress f32 view)");
            const unsigned grid = (unsigned)std::min<int64_t>((len + 255) / 256, (int64_t)sm_count() * 16);
            if (sp.dtype == MMC_F64) progress_cast_kernel<double><<<grid, 256, 0, stream>>>(reinterpret_cast<const double *>(d_sample), d_f32, len);
            else progress_cast_kernel<unsigned long long><<<grid, 256, 0, stream>>>(reinterpret_cast<const unsigned long long *>(d_sample), d_f32, len);
            f32 = d_f32;
        }
        std::vector<float> rhat((size_t)sp.dim), ess((size_t)sp.dim);
        if ((rc = mmc_split_rhat_ess_dev(f32, sp.chains, n_collect, sp.dim, rhat.data(), ess.data(), stream))) { cleanup(); return rc; }
        if ((rc = mmc_basic_stats_of(ess.data(), sp.dim, &stats->ess))) { cleanup(); return rc; }
        if ((rc = mmc_basic_stats_of(rhat.data(), sp.dim, &stats->rhat))) { cleanup(); return rc; }
    }
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return fail_cuda(e, "run_progress sync");
    cleanup();
    return MMC_OK;
}

}  // namespace mmc
