// NUTS C ABI (see include/minimcmc.h) — host side of K4.
#include <climits>
#include <vector>

#include "mmc_nuts_group_inst.cuh"
#include "mmc_progress.cuh"

using namespace mmc;

struct mmc_nuts {
    mmc_target_desc target{};
    int64_t chains = 0;
    int32_t dim = 0;
    double target_accept = 0.8;
    int32_t scalar_dtype = MMC_F64;
    int32_t max_depth = 10;
    int64_t chain_offset = 0;
    uint64_t seed = 0;
    int32_t exact = 0;
    int64_t out_pitch = 0;     // *_dev runs: draws per chain row of the caller's tensor (0 = n_collect)
    int64_t adapt_until = -1;  // -1: the reference's rule (m <= n_discard of the call)
    int32_t resume = 0;        // 1: the next run continues the chains without init_chain
    int32_t layout = 0;        // lanes per chain requested: 0 = auto, 32 = one chain per warp, else the group kernel's G
    int32_t lanes_used = 0;    // lanes per chain of the last launch
    float *d_pos = nullptr;
    double *d_state = nullptr;            // [chains, 5]
    unsigned long long *d_counters = nullptr;  // [8 + 32]
    float *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    int *d_flags = nullptr;    // group kernel: completed slices per group of chains
    size_t flags_bytes = 0;
    int64_t slice_steps = -1;  // -1: automatic (a sixteenth of the launch's iterations, at least 8), 0: off
    int32_t regroup = -1;      // group kernel: re-form the warps' chain groups by step size between phases (1 = on; -1 / 0 = off)
    int *d_perm = nullptr;     // [chains] chain order of the current phase, [chains] bin of every chain, [2 kRegroupBins] counts / offsets
    size_t perm_bytes = 0;
    cudaStream_t stream = nullptr;
    float *d_out = nullptr;
    size_t d_out_bytes = 0;
    double *d_tape[3] = {nullptr, nullptr, nullptr};
    size_t d_tape_bytes[3] = {0, 0, 0};
    double *trace_dev = nullptr;  // caller-owned per-transition trace of the next runs (mmc_nuts_set_trace_dev)
    int64_t trace_pitch = 0;
};

namespace {

constexpr int kCounters = 8 + 32;

template <class T>
int grow(T **ptr, size_t *cap, size_t need) {
    if (*cap >= need) return MMC_OK;
    if (*ptr) MMC_CUDA(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    MMC_CUDA(cudaMalloc((void **)ptr, need));
    *cap = need;
    return MMC_OK;
}

// one launch (or grid / scratch query) of the tree kernel the handle's settings select: a built-in target on the group
// or warp layout, or the launcher a custom target registered (mmc_register_nuts_target; warp layout)
int nuts_launch(const NutsLaunch &L, int exact, bool group, const NutsParams &p, int64_t *grid, size_t *scratch, bool query,
                cudaStream_t s) {
    if (L.target.kind >= MMC_T_CUSTOM_BASE) {
        CustomTargetEntry e;
        MMC_REQUIRE(custom_target_get(L.target.kind, &e) && e.nuts, "custom target kind %d is not registered for NUTS", L.target.kind);
        return e.nuts(&p, L.scalar_f64 ? 1 : 0, L.replay ? 1 : 0, exact, L.target.params, L.sm_count, grid, scratch, query ? 1 : 0, s);
    }
    auto dispatch = group ? (exact ? nuts_group_dispatch_exact : nuts_group_dispatch_fast)
                          : (exact ? nuts_dispatch_exact : nuts_dispatch_fast);
    return dispatch(L, p, grid, scratch, query, s);
}

// ---- regrouping of the group kernel's warps (lock-step efficiency)
// The chains of a warp advance in lock step, so a warp is as slow as its deepest tree; tree depth is governed by the
// chain's adapted step size (oracle study in DESIGN.md: corr(log eps, leapfrogs per chain) = -0.92; lock-step efficiency
// 0.73 with index-ordered groups, 0.78-0.86 with groups of similar step size).  Between the phases of a run a counting
// sort by log2(eps) (2,048 bins of 1/64 octave) builds the chain order of the next launch.  Draws do not depend on the
// grouping (Philox is keyed by the global chain id), so the order inside a bin is left to the atomics.
constexpr int kRegroupBins = 2048;

__global__ void nuts_regroup_hist(const double *state, int64_t chains, int64_t adapt_until, int *hist, int *bin_of) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chains) return;
    // the step size of the next transition: epsilon while dual averaging runs, epsilon_bar afterwards (src/nuts.rs:681-690)
    const double m = state[c * 5 + 4];
    const double eps = (m >= (double)adapt_until) ? state[c * 5 + 1] : state[c * 5 + 0];
    int bin = 0;
    if (eps > 0.0 && eps < 1e30) {
        bin = (int)((log2f((float)eps) + 24.0f) * (kRegroupBins / 32.0f));
        bin = bin < 0 ? 0 : (bin >= kRegroupBins ? kRegroupBins - 1 : bin);
    }
    bin_of[c] = bin;
    atomicAdd(&hist[bin], 1);
}

__global__ void nuts_regroup_scan(const int *hist, int *offs) {  // one block of 1,024 threads, 2 bins each
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const int a = hist[2 * t], b = hist[2 * t + 1];
    part[t] = a + b;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const int base = part[t] - (a + b);
    offs[2 * t] = base;
    offs[2 * t + 1] = base + a;
}

__global__ void nuts_regroup_scatter(const int *bin_of, int64_t chains, int *offs, int *perm) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chains) return;
    perm[atomicAdd(&offs[bin_of[c]], 1)] = (int)c;
}

__global__ void nuts_merge_test_kernel(const uint64_t *k53, const uint32_t *num, const uint32_t *den, int64_t n, uint8_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = u53_below_ratio(k53[i], num[i], den[i]) ? 1 : 0;
}

}  // namespace

extern "C" {

// Debug entry: evaluates the tree-merge acceptance test "k 2^-53 < num / den" (src/nuts.rs:910-911 with a native 53-bit
// draw) on the device for n triples; tests/test_gpu_nuts.py checks it against exact integer arithmetic up to the tree
// sizes of max_depth = 16.
int mmc_debug_nuts_merge_test(const uint64_t *k53_host, const uint32_t *num_host, const uint32_t *den_host, int64_t n,
                              uint8_t *out_host) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(k53_host && num_host && den_host && out_host && n > 0, "mmc_debug_nuts_merge_test: bad arguments");
    unsigned char *d = nullptr;
    MMC_CUDA(cudaMalloc((void **)&d, (size_t)n * 17));
    uint64_t *dk = reinterpret_cast<uint64_t *>(d);
    uint32_t *dn = reinterpret_cast<uint32_t *>(d + n * 8), *dd = dn + n;
    uint8_t *dout = d + n * 16;
    cudaMemcpy(dk, k53_host, (size_t)n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dn, num_host, (size_t)n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dd, den_host, (size_t)n * 4, cudaMemcpyHostToDevice);
    nuts_merge_test_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dk, dn, dd, n, dout);
    const cudaError_t e = cudaMemcpy(out_host, dout, (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "mmc_debug_nuts_merge_test", __FILE__, __LINE__);
    return MMC_OK;
}

int mmc_nuts_create(mmc_nuts **out, const mmc_target_desc *target, const float *init_host, int64_t chains,
                    int32_t dim, double target_accept_p, int32_t scalar_dtype, int32_t max_depth) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(out && target && init_host && chains > 0 && dim > 0, "mmc_nuts_create: bad arguments");
    MMC_REQUIRE(target->dim == dim, "target dim %d != dim %d", target->dim, dim);
    MMC_REQUIRE(scalar_dtype == MMC_F32 || scalar_dtype == MMC_F64, "scalar dtype must be MMC_F32 or MMC_F64");
    if (target->kind >= MMC_T_CUSTOM_BASE) {
        CustomTargetEntry e;
        const char *nm = "";
        MMC_REQUIRE(custom_target_get(target->kind, &e, &nm) && e.nuts, "custom target kind %d is not registered for NUTS", target->kind);
        MMC_REQUIRE(e.dim == dim, "custom target '%s' has dim %d, got %d", nm, e.dim, dim);
    }
    if (max_depth <= 0) max_depth = 10;
    MMC_REQUIRE(max_depth <= 16, "max_depth %d > 16", max_depth);
    mmc_nuts *h = new mmc_nuts();
    h->target = *target;
    h->chains = chains;
    h->dim = dim;
    h->target_accept = target_accept_p;
    h->scalar_dtype = scalar_dtype;
    h->max_depth = max_depth;
    auto fail = [&](cudaError_t e, const char *what) {
        int code = cuda_fail(e, what, __FILE__, __LINE__);
        mmc_nuts_destroy(h);
        return code;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    const size_t bytes = (size_t)chains * dim * sizeof(float);
    if ((e = cudaMalloc((void **)&h->d_pos, bytes)) != cudaSuccess) return fail(e, "cudaMalloc(positions)");
    if ((e = cudaMemcpy(h->d_pos, init_host, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "memcpy");
    // NUTSChain::new, src/nuts.rs:410-434: epsilon = -1 (unset), epsilon_bar = 1, h_bar = 0, mu = ln 10, m = 0
    std::vector<double> st((size_t)chains * 5);
    for (int64_t c = 0; c < chains; ++c) {
        st[c * 5 + 0] = -1.0; st[c * 5 + 1] = 1.0; st[c * 5 + 2] = 0.0; st[c * 5 + 3] = std::log(10.0); st[c * 5 + 4] = 0.0;
    }
    if ((e = cudaMalloc((void **)&h->d_state, st.size() * 8)) != cudaSuccess) return fail(e, "cudaMalloc(state)");
    if ((e = cudaMemcpy(h->d_state, st.data(), st.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "memcpy");
    if ((e = cudaMalloc((void **)&h->d_counters, kCounters * 8)) != cudaSuccess) return fail(e, "cudaMalloc(counters)");
    if ((e = cudaMemset(h->d_counters, 0, kCounters * 8)) != cudaSuccess) return fail(e, "memset");
    *out = h;
    return MMC_OK;
}

int mmc_register_nuts_target(const char *name, int32_t dim, mmc_nuts_launch_fn fn) {
    MMC_REQUIRE(name && fn && dim > 0 && dim <= 128, "mmc_register_nuts_target: bad arguments (dim <= 128)");
    return custom_target_register(name, dim, nullptr, fn, nullptr);
}

int mmc_nuts_set_seed(mmc_nuts *h, uint64_t seed) {
    MMC_REQUIRE(h, "null handle");
    h->seed = seed;
    return MMC_OK;
}

int mmc_nuts_set_chain_offset(mmc_nuts *h, int64_t offset) {
    MMC_REQUIRE(h && offset >= 0, "bad chain offset");
    h->chain_offset = offset;
    return MMC_OK;
}

int mmc_nuts_set_exact(mmc_nuts *h, int32_t exact) {
    MMC_REQUIRE(h, "null handle");
    h->exact = exact ? 1 : 0;
    return MMC_OK;
}

int mmc_nuts_set_layout(mmc_nuts *h, int32_t lanes_per_chain) {
    MMC_REQUIRE(h, "null handle");
    const int avail = nuts_group_lanes(h->target);
    MMC_REQUIRE(lanes_per_chain == kNutsLayoutAuto || lanes_per_chain == kNutsLayoutWarp || lanes_per_chain == avail,
                "mmc_nuts_set_layout: %d lanes per chain is not compiled in for this target (0 = auto, 32 = warp per chain%s%d)",
                lanes_per_chain, avail ? ", group = " : "; no group layout, ", avail);
    h->layout = lanes_per_chain;
    return MMC_OK;
}

int mmc_nuts_set_slicing(mmc_nuts *h, int64_t slice_steps) {
    MMC_REQUIRE(h && slice_steps >= -1, "mmc_nuts_set_slicing: bad arguments");
    h->slice_steps = slice_steps;
    return MMC_OK;
}

int mmc_nuts_set_regroup(mmc_nuts *h, int32_t mode) {
    MMC_REQUIRE(h && mode >= -1 && mode <= 1, "mmc_nuts_set_regroup: mode must be -1 (auto), 0 (off) or 1 (on)");
    h->regroup = mode;
    return MMC_OK;
}

int mmc_nuts_get_layout(mmc_nuts *h, int32_t *lanes_per_chain) {
    MMC_REQUIRE(h && lanes_per_chain, "mmc_nuts_get_layout: bad arguments");
    *lanes_per_chain = h->lanes_used;
    return MMC_OK;
}

int mmc_nuts_set_out_pitch(mmc_nuts *h, int64_t pitch_steps) {
    MMC_REQUIRE(h && pitch_steps >= 0, "mmc_nuts_set_out_pitch: bad arguments");
    h->out_pitch = pitch_steps;
    return MMC_OK;
}

int mmc_nuts_set_continuation(mmc_nuts *h, int64_t adapt_until, int32_t resume) {
    MMC_REQUIRE(h && adapt_until >= -1, "mmc_nuts_set_continuation: bad arguments");
    h->adapt_until = adapt_until;
    h->resume = resume ? 1 : 0;
    return MMC_OK;
}

int mmc_nuts_run_dev(mmc_nuts *h, int64_t n_collect, int64_t n_discard, int32_t progress, float *out_dev,
                     const mmc_replay_nuts *rp, void *stream) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0, "mmc_nuts_run_dev: bad arguments");
    MMC_REQUIRE(out_dev || n_collect == 0, "mmc_nuts_run_dev: out is null");
    MMC_REQUIRE(h->out_pitch == 0 || h->out_pitch >= n_collect, "mmc_nuts_run_dev: out pitch %lld < n_collect", (long long)h->out_pitch);
    const bool replay = rp && rp->normals && rp->exps && rp->unifs;
    MMC_REQUIRE(!rp || replay, "NUTS replay needs normals, exps and unifs tapes");
    cudaStream_t s = (cudaStream_t)stream;
    NutsParams p{};
    p.positions = h->d_pos;
    p.out = out_dev;
    p.state = h->d_state;
    if (replay) {
        p.normals = rp->normals; p.exps = rp->exps; p.unifs = rp->unifs;
        p.cap_normals = rp->cap_normals; p.cap_exps = rp->cap_exps; p.cap_unifs = rp->cap_unifs;
    }
    p.counters = h->d_counters;
    p.chains = h->chains;
    p.chain_offset = h->chain_offset;
    p.n_collect = n_collect;
    p.n_discard = n_discard;
    p.out_pitch = h->out_pitch > 0 ? h->out_pitch : n_collect;
    p.adapt_until = h->adapt_until >= 0 ? h->adapt_until : n_discard;
    p.resume = h->resume;
    p.progress = progress ? 1 : 0;
    p.max_depth = h->max_depth;
    p.D = h->dim;
    p.target_accept = h->target_accept;
    p.key = seed_key(h->seed);
    p.it_lo = 0;
    p.it_hi = -1;  // the whole run unless the phased launches below narrow it
    p.trace = h->trace_dev;
    p.trace_pitch = h->trace_pitch;
    MMC_REQUIRE(!p.trace || p.trace_pitch >= n_collect + n_discard, "mmc_nuts_run_dev: trace pitch %lld < %lld iterations",
                (long long)p.trace_pitch, (long long)(n_collect + n_discard));
    NutsLaunch L{h->target, h->scalar_dtype == MMC_F64, replay, sm_count()};
    // several chains per warp wherever that layout is compiled in for the target, unless the caller pins the layout
    const int group_lanes = nuts_group_lanes(h->target);
    const bool group = h->layout == kNutsLayoutAuto ? group_lanes != 0 : h->layout != kNutsLayoutWarp;
    h->lanes_used = group ? group_lanes : 32;
    const int exact = h->exact;
    auto dispatch = [&](const NutsLaunch &LL, const NutsParams &pp, int64_t *g, size_t *sc, bool query, cudaStream_t st) {
        return nuts_launch(LL, exact, group, pp, g, sc, query, st);
    };
    int64_t grid = 0;
    size_t scratch_floats = 0;
    int rc = dispatch(L, p, &grid, &scratch_floats, true, s);
    if (rc) return rc;
    if ((rc = grow(&h->d_scratch, &h->scratch_bytes, scratch_floats * sizeof(float) + 16))) return rc;
    p.scratch = h->d_scratch;
    if (!(group && !replay)) {
        MMC_CUDA(cudaMemsetAsync(h->d_counters, 0, 8, s));  // work-item ticket
        return dispatch(L, p, &grid, &scratch_floats, false, s);
    }
    // Native runs of the group kernel.  The run is cut into phases (one launch each) at iterations 32, 96, 224 and at the
    // end of the burn-in; from the second phase on the warps' groups are re-formed from chains of similar step size.
    // Inside a launch the work is dispensed in slices (replay keeps single launches: its cursors live in registers).
    const int chains_per_warp = 32 / group_lanes;
    const size_t n_groups = (size_t)(h->chains + chains_per_warp - 1) / chains_per_warp;
    if ((rc = grow(&h->d_flags, &h->flags_bytes, n_groups * sizeof(int)))) return rc;
    const int64_t total = n_collect + n_discard, first = (progress || h->resume) ? 0 : 1;
    const bool regroup = h->regroup > 0 && h->chains <= INT32_MAX;  // opt-in: measured neutral on C5
    std::vector<int64_t> bounds{first};
    if (regroup) {
        for (int64_t b : {(int64_t)32, (int64_t)96, (int64_t)224, n_discard})
            if (b >= bounds.back() + 16 && b + 16 <= total) bounds.push_back(b);
        const size_t need = ((size_t)h->chains * 2 + 2 * kRegroupBins) * sizeof(int);
        if (bounds.size() > 1 && (rc = grow(&h->d_perm, &h->perm_bytes, need))) return rc;
    }
    bounds.push_back(total > first ? total : first);
    for (size_t w = 0; w + 1 < bounds.size(); ++w) {
        const int64_t lo = bounds[w], hi = bounds[w + 1];
        if (w > 0) {
            if (hi <= lo) continue;
            int *perm = h->d_perm, *bin_of = perm + h->chains, *hist = bin_of + h->chains, *offs = hist + kRegroupBins;
            const unsigned blocks = (unsigned)((h->chains + 255) / 256);
            MMC_CUDA(cudaMemsetAsync(hist, 0, kRegroupBins * sizeof(int), s));
            nuts_regroup_hist<<<blocks, 256, 0, s>>>(h->d_state, h->chains, p.adapt_until, hist, bin_of);
            nuts_regroup_scan<<<1, 1024, 0, s>>>(hist, offs);
            nuts_regroup_scatter<<<blocks, 256, 0, s>>>(bin_of, h->chains, offs, perm);
            MMC_CUDA(cudaGetLastError());
            p.perm = perm;
        }
        const int64_t n_iter = hi - lo;
        int64_t slice = h->slice_steps >= 0 ? h->slice_steps : (n_iter + 15) / 16;
        if (slice > 0 && slice < 8) slice = 8;
        p.slice_steps = slice;
        p.flags = h->d_flags;
        p.it_lo = lo;
        p.it_hi = w + 2 == bounds.size() ? -1 : hi;
        MMC_CUDA(cudaMemsetAsync(h->d_flags, 0, n_groups * sizeof(int), s));
        MMC_CUDA(cudaMemsetAsync(h->d_counters, 0, 8, s));  // work-item ticket
        if ((rc = dispatch(L, p, &grid, &scratch_floats, false, s))) return rc;
    }
    return MMC_OK;
}

int mmc_nuts_run(mmc_nuts *h, int64_t n_collect, int64_t n_discard, int32_t progress, float *out_host,
                 const mmc_replay_nuts *replay) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0 && (out_host || n_collect == 0), "mmc_nuts_run: bad arguments");
    MMC_REQUIRE(h->out_pitch == 0, "mmc_nuts_run: an output pitch only applies to mmc_nuts_run_dev");
    const size_t out_bytes = (size_t)h->chains * n_collect * h->dim * sizeof(float);
    int rc = grow(&h->d_out, &h->d_out_bytes, out_bytes ? out_bytes : 16);
    if (rc) return rc;
    mmc_replay_nuts dev_rp{};
    const mmc_replay_nuts *rp = nullptr;
    if (replay) {
        MMC_REQUIRE(replay->normals && replay->exps && replay->unifs, "NUTS replay needs normals, exps and unifs");
        const double *src[3] = {replay->normals, replay->exps, replay->unifs};
        const int64_t cap[3] = {replay->cap_normals, replay->cap_exps, replay->cap_unifs};
        for (int i = 0; i < 3; ++i) {
            const size_t b = (size_t)h->chains * cap[i] * 8;
            if ((rc = grow(&h->d_tape[i], &h->d_tape_bytes[i], b ? b : 8))) return rc;
            MMC_CUDA(cudaMemcpyAsync(h->d_tape[i], src[i], b, cudaMemcpyHostToDevice, h->stream));
        }
        dev_rp = *replay;
        dev_rp.normals = h->d_tape[0]; dev_rp.exps = h->d_tape[1]; dev_rp.unifs = h->d_tape[2];
        rp = &dev_rp;
    }
    rc = mmc_nuts_run_dev(h, n_collect, n_discard, progress, h->d_out, rp, h->stream);
    if (rc) return rc;
    if (out_bytes) MMC_CUDA(cudaMemcpyAsync(out_host, h->d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    MMC_CUDA(cudaStreamSynchronize(h->stream));
    return MMC_OK;
}

int mmc_nuts_run_progress(mmc_nuts *h, int64_t n_collect, int64_t n_discard, float *out_host, int64_t block, mmc_progress_fn cb,
                          void *user, mmc_run_stats *stats) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0 && (out_host || n_collect == 0), "mmc_nuts_run_progress: bad arguments");
    MMC_REQUIRE(h->out_pitch == 0, "mmc_nuts_run_progress: an output pitch is set on this handle");
    ProgressSpec sp{h->chains, h->dim, MMC_F32, MMC_TRACK_PER_CHAIN, true, h->d_pos};
    const int64_t saved_until = h->adapt_until;
    const int32_t saved_resume = h->resume;
    auto run_block = [&](int64_t k, void *dst, int64_t pitch, bool first) {
        // the reference compares the chain's absolute step count m with this call's n_discard (src/nuts.rs:681); later
        // blocks continue without init_chain so that the blocks reproduce the single-launch run
        h->out_pitch = pitch;
        h->adapt_until = n_discard;
        h->resume = first ? 0 : 1;
        const int rc = mmc_nuts_run_dev(h, k, 0, 1, static_cast<float *>(dst), nullptr, h->stream);
        h->out_pitch = 0;
        h->adapt_until = saved_until;
        h->resume = saved_resume;
        return rc;
    };
    auto discard = [&](int64_t k) { return mmc_nuts_run_dev(h, 0, k, 1, nullptr, nullptr, h->stream); };
    return run_progress_blocks(sp, n_collect, n_discard, out_host, block, cb, user, stats, h->stream, run_block, discard);
}

int mmc_nuts_set_trace_dev(mmc_nuts *h, double *trace_dev, int64_t pitch_steps) {
    MMC_REQUIRE(h && pitch_steps >= 0 && (!trace_dev || pitch_steps > 0), "mmc_nuts_set_trace_dev: bad arguments");
    h->trace_dev = trace_dev;
    h->trace_pitch = pitch_steps;
    return MMC_OK;
}

int mmc_nuts_set_state(mmc_nuts *h, const double *state_host) {
    MMC_REQUIRE(h && state_host, "mmc_nuts_set_state: bad arguments");
    MMC_CUDA(cudaDeviceSynchronize());
    MMC_CUDA(cudaMemcpy(h->d_state, state_host, (size_t)h->chains * 5 * 8, cudaMemcpyHostToDevice));
    return MMC_OK;
}

int mmc_nuts_set_positions(mmc_nuts *h, const float *positions_host) {
    MMC_REQUIRE(h && positions_host, "mmc_nuts_set_positions: bad arguments");
    MMC_CUDA(cudaDeviceSynchronize());
    MMC_CUDA(cudaMemcpy(h->d_pos, positions_host, (size_t)h->chains * h->dim * sizeof(float), cudaMemcpyHostToDevice));
    return MMC_OK;
}

// Debug / known-answer entry: ONE build_tree(position, momentum, grad, logu, v, j, epsilon, joint_0) per chain
// (src/nuts.rs:764-946, test_build_tree :1057-1121) on the kernels and lane layout the handle would run, with the
// uniforms read from a per-chain tape.  The positions of the handle are the input positions and are left unchanged.
int mmc_nuts_build_tree(mmc_nuts *h, const float *mom_host, const float *grad_host, const double *scal_host, int32_t j,
                        const double *unifs_host, int64_t cap_unifs, float *out_vec_host, double *out_scal_host) {
    MMC_REQUIRE(h && mom_host && grad_host && scal_host && unifs_host && out_vec_host && out_scal_host && cap_unifs > 0,
                "mmc_nuts_build_tree: bad arguments");
    MMC_REQUIRE(j >= 0 && j <= h->max_depth, "mmc_nuts_build_tree: depth %d outside [0, max_depth = %d]", j, h->max_depth);
    const size_t nv = (size_t)h->chains * h->dim;
    float *d_vec = nullptr;   // mom, grad, out_vec[5]
    double *d_scal = nullptr; // scal[4], out_scal[6], unifs[cap]
    cudaStream_t s = h->stream;
    int rc = MMC_OK;
    auto done = [&](int code) {
        cudaFree(d_vec);
        cudaFree(d_scal);
        return code;
    };
    MMC_CUDA(cudaMalloc((void **)&d_vec, nv * 7 * sizeof(float)));
    if (cudaMalloc((void **)&d_scal, (size_t)h->chains * (10 + cap_unifs) * 8) != cudaSuccess) return done(MMC_ERR_CUDA);
    double *d_si = d_scal, *d_so = d_scal + h->chains * 4, *d_un = d_scal + h->chains * 10;
    cudaMemcpyAsync(d_vec, mom_host, nv * 4, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_vec + nv, grad_host, nv * 4, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_si, scal_host, (size_t)h->chains * 4 * 8, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_un, unifs_host, (size_t)h->chains * cap_unifs * 8, cudaMemcpyHostToDevice, s);
    NutsParams p{};
    p.positions = h->d_pos;
    p.state = h->d_state;
    p.normals = d_un; p.exps = d_un; p.unifs = d_un;  // only the uniform tape is read
    p.cap_normals = p.cap_exps = p.cap_unifs = cap_unifs;
    p.counters = h->d_counters;
    p.chains = h->chains;
    p.chain_offset = h->chain_offset;
    p.max_depth = h->max_depth;
    p.D = h->dim;
    p.target_accept = h->target_accept;
    p.key = seed_key(h->seed);
    p.it_hi = -1;
    p.progress = 1;
    p.tree_mom = d_vec;
    p.tree_grad = d_vec + nv;
    p.tree_scal = d_si;
    p.tree_out_vec = d_vec + 2 * nv;
    p.tree_out_scal = d_so;
    p.tree_j = j;
    NutsLaunch L{h->target, h->scalar_dtype == MMC_F64, true, sm_count()};
    const int group_lanes = nuts_group_lanes(h->target);
    const bool group = h->layout == kNutsLayoutAuto ? group_lanes != 0 : h->layout != kNutsLayoutWarp;
    h->lanes_used = group ? group_lanes : 32;
    const int exact = h->exact;
    auto dispatch = [&](const NutsLaunch &LL, const NutsParams &pp, int64_t *g, size_t *sc, bool query, cudaStream_t st) {
        return nuts_launch(LL, exact, group, pp, g, sc, query, st);
    };
    int64_t grid = 0;
    size_t scratch_floats = 0;
    if ((rc = dispatch(L, p, &grid, &scratch_floats, true, s))) return done(rc);
    if ((rc = grow(&h->d_scratch, &h->scratch_bytes, scratch_floats * sizeof(float) + 16))) return done(rc);
    p.scratch = h->d_scratch;
    cudaMemsetAsync(h->d_counters, 0, 8, s);
    if ((rc = dispatch(L, p, &grid, &scratch_floats, false, s))) return done(rc);
    cudaMemcpyAsync(out_vec_host, d_vec + 2 * nv, nv * 5 * 4, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(out_scal_host, d_so, (size_t)h->chains * 6 * 8, cudaMemcpyDeviceToHost, s);
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return done(cuda_fail(e, "mmc_nuts_build_tree", __FILE__, __LINE__));
    return done(MMC_OK);
}

int mmc_nuts_get_state(mmc_nuts *h, double *state_host) {
    MMC_REQUIRE(h && state_host, "mmc_nuts_get_state: bad arguments");
    MMC_CUDA(cudaDeviceSynchronize());
    MMC_CUDA(cudaMemcpy(state_host, h->d_state, (size_t)h->chains * 5 * 8, cudaMemcpyDeviceToHost));
    return MMC_OK;
}

int mmc_nuts_get_positions(mmc_nuts *h, float *positions_host) {
    MMC_REQUIRE(h && positions_host, "mmc_nuts_get_positions: bad arguments");
    MMC_CUDA(cudaDeviceSynchronize());
    MMC_CUDA(cudaMemcpy(positions_host, h->d_pos, (size_t)h->chains * h->dim * sizeof(float), cudaMemcpyDeviceToHost));
    return MMC_OK;
}

int mmc_nuts_get_counters(mmc_nuts *h, int64_t *n_grad, int64_t *n_transitions, int64_t *depth_hist, int32_t hist_len) {
    MMC_REQUIRE(h, "null handle");
    unsigned long long c[kCounters];
    MMC_CUDA(cudaDeviceSynchronize());
    MMC_CUDA(cudaMemcpy(c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
    if (n_grad) *n_grad = (int64_t)c[1];
    if (n_transitions) *n_transitions = (int64_t)c[2];
    for (int i = 0; depth_hist && i < hist_len; ++i) depth_hist[i] = i < 32 ? (int64_t)c[8 + i] : 0;
    return MMC_OK;
}

void mmc_nuts_destroy(mmc_nuts *h) {
    if (!h) return;
    cudaFree(h->d_pos);
    cudaFree(h->d_state);
    cudaFree(h->d_counters);
    cudaFree(h->d_scratch);
    cudaFree(h->d_flags);
    cudaFree(h->d_perm);
    cudaFree(h->d_out);
    for (auto p : h->d_tape) cudaFree(p);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

}  // extern "C"
