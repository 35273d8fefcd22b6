// Metropolis-Hastings C ABI (see include/minimcmc.h) — host side of K1.
// Compiled with -fmad=false (see mmc_mh.cuh).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "mmc_mh.cuh"
#include "mmc_progress.cuh"

using namespace mmc;

struct mmc_mh {
    mmc_target_desc target{};
    mmc_proposal_desc proposal{};
    int64_t chains = 0;
    int32_t dim = 0;
    int32_t dtype = MMC_F64;
    int64_t chain_offset = 0;
    int64_t step = 0;  // transitions since the last seed()
    uint64_t seed = 0;
    int32_t accept_mode = 1;
    int64_t out_pitch = 0;  // *_dev runs: draws per chain row of the caller's tensor (0 = n_collect)
    int32_t tile_u8 = 256, tile_u16 = 128;  // staging-tile steps (tuned on B200)
    void *d_state = nullptr;
    // Poisson tables
    int32_t table_len = 0;
    double *d_lnfact = nullptr;
    uint4 *d_lim = nullptr;  // [table_len][2]: (thr >> 38, thr_lo low 32, thr_lo high 6, 0)
    int32_t *d_error = nullptr;
    double ln_lambda = 0, ln_half = 0;
    // host-API staging
    cudaStream_t stream = nullptr;
    void *d_out = nullptr;
    size_t d_out_bytes = 0;
    void *d_replay[3] = {nullptr, nullptr, nullptr};
    size_t d_replay_bytes[3] = {0, 0, 0};
    double *d_trace = nullptr;
    size_t d_trace_bytes = 0;
    // compact device->host pipeline (Poisson): double-buffered device + pinned host staging
    void *d_compact[2] = {nullptr, nullptr};
    void *d_compact_only[2] = {nullptr, nullptr};   // mmc_mh_run_compact: device blocks copied straight to the caller
    size_t compact_only_bytes = 0;
    void *h_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    cudaStream_t pipe_stream[2] = {nullptr, nullptr};
    cudaEvent_t pipe_event[2] = {nullptr, nullptr};
};

namespace mmc {
void widen_to_u64(const void *src, int elem_bytes, uint64_t *dst, size_t n);
}

namespace {

size_t elem_size(int32_t dtype) { return dtype == MMC_F32 ? 4 : 8; }

// integer-state targets driven by the threshold tables: Poisson, Categorical and tabulated log-probabilities
bool is_int_target(const mmc_mh *h) {
    return h->target.kind == MMC_T_POISSON || h->target.kind == MMC_T_CATEGORICAL || h->target.kind == MMC_T_TABULATED;
}

int grow(void **ptr, size_t *cap, size_t need) {
    if (*cap >= need) return MMC_OK;
    if (*ptr) MMC_CUDA(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    MMC_CUDA(cudaMalloc(ptr, need));
    *cap = need;
    return MMC_OK;
}

// Smallest m in [0, 2^53] with !(r > ln(m * 2^-53)): the accept set {u : r > ln u} over the 53-bit
// uniform grid is [0, m) because ln is monotone.  Evaluated with the host libm, the same ln the
// reference's CPU path calls.
uint64_t accept_threshold(double r) {
    if (!(r > -INFINITY)) return 0;  // r = -inf or NaN: never accepted (even ln(0) = -inf fails '>')
    const uint64_t top = 1ULL << 53;
    auto pred = [&](uint64_t m) { return r > std::log((double)m * 0x1p-53); };
    if (pred(top - 1)) return top;
    uint64_t lo = 0, hi = top - 1;  // pred(lo) true, pred(hi) false
    while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (pred(mid)) lo = mid; else hi = mid;
    }
    return hi;
}

int upload_int_tables(mmc_mh *h, const std::vector<double> &lp, const std::vector<double> &lnfact);

int build_poisson_tables(mmc_mh *h) {
    const double lambda = h->target.params[0];
    MMC_REQUIRE(lambda > 0.0, "Poisson target needs lambda > 0");
    // generous range: mean + 40 sd + slack, bounded by the u16 staging tile
    int64_t len = (int64_t)std::ceil(lambda + 40.0 * std::sqrt(lambda) + 64.0);
    if (len < 256) len = 256;
    if (len > 4096) len = 4096;
    h->table_len = (int32_t)len;
    h->ln_lambda = std::log(lambda);
    std::vector<double> lnfact(len), lp(len);
    // ln_factorial, examples/poisson_mh.rs:79-89: 0 for k < 2, else sum_{i=1..k} ln(i) in that order.
    double acc = 0.0;
    for (int64_t k = 0; k < len; ++k) {
        if (k >= 1) acc += std::log((double)k);
        lnfact[k] = k < 2 ? 0.0 : acc;
        lp[k] = -lambda + (double)k * h->ln_lambda - lnfact[k];
    }
    return upload_int_tables(h, lp, lnfact);
}

// Categorical::new / logp, src/distributions.rs:431-468: probs normalised by their left-fold sum, logp = ln(p_k) for
// k < K and -inf beyond.  The table carries 16 unreachable states past K (the kernel's overflow check is conservative
// by 7 inside an octet).
int build_categorical_tables(mmc_mh *h, const double *probs, int32_t n) {
    double sum = 0.0;
    for (int32_t k = 0; k < n; ++k) sum = sum + probs[k];
    const int64_t len = (int64_t)n + 16;
    h->table_len = (int32_t)len;
    std::vector<double> lnfact(len, 0.0), lp(len, -INFINITY);
    for (int32_t k = 0; k < n; ++k) lp[k] = std::log(probs[k] / sum);
    return upload_int_tables(h, lp, lnfact);
}

// Host-built accept thresholds of the +-1 nonnegative random walk (examples/poisson_mh.rs:28-77) for any integer
// target given as a table of log-probabilities.
int upload_int_tables(mmc_mh *h, const std::vector<double> &lp, const std::vector<double> &lnfact) {
    const int64_t len = (int64_t)lp.size();
    h->ln_half = std::log(0.5);
    // table entry [k][dir] = (thr >> 38, thr & (2^38 - 1)):  u53 < thr  <=>  u15 < thr_hi || (u15 == thr_hi && u38 < thr_lo)
    std::vector<uint4> lim(2 * len, make_uint4(0u, 0u, 0u, 0u));
    auto encode = [](uint64_t thr) {
        const uint64_t lo = thr & ((1ULL << 38) - 1);
        return make_uint4((uint32_t)(thr >> 38), (uint32_t)lo, (uint32_t)(lo >> 32), 0u);
    };
    const bool reflect = h->proposal.kind == MMC_Q_REFLECT_RW;
    for (int64_t k = 0; k + 1 < len; ++k) {
        if (reflect) {
            // symmetric +-1 walk clamped to the support (tests/metrohast_poisson_test.rs:63-80,193-207): q_f = q_b = ln 1/2.
            // A clamped move proposes the current state, so leaving its threshold at 0 ("never accept") gives the same chain.
            lim[2 * k + 1] = encode(accept_threshold((lp[k + 1] + h->ln_half) - (lp[k] + h->ln_half)));
            if (k >= 1) lim[2 * k] = encode(accept_threshold((lp[k - 1] + h->ln_half) - (lp[k] + h->ln_half)));
            continue;
        }
        // x = k -> y = k + 1 : q_f = (k == 0 ? 0 : ln 1/2), q_b = ln 1/2
        const double qf = k == 0 ? 0.0 : h->ln_half, qb = h->ln_half;
        lim[2 * k + 1] = encode(accept_threshold((lp[k + 1] + qb) - (lp[k] + qf)));
        if (k >= 1) {
            // x = k -> y = k - 1 : q_f = ln 1/2, q_b = (y == 0 ? 0 : ln 1/2)
            const double qb2 = (k - 1 == 0) ? 0.0 : h->ln_half;
            lim[2 * k] = encode(accept_threshold((lp[k - 1] + qb2) - (lp[k] + h->ln_half)));
        }
    }
    MMC_CUDA(cudaMalloc(&h->d_lnfact, len * sizeof(double)));
    MMC_CUDA(cudaMalloc(&h->d_lim, 2 * len * sizeof(uint4)));
    MMC_CUDA(cudaMalloc(&h->d_error, sizeof(int32_t)));
    MMC_CUDA(cudaMemcpy(h->d_lnfact, lnfact.data(), len * sizeof(double), cudaMemcpyHostToDevice));
    MMC_CUDA(cudaMemcpy(h->d_lim, lim.data(), 2 * len * sizeof(uint4), cudaMemcpyHostToDevice));
    MMC_CUDA(cudaMemset(h->d_error, 0, sizeof(int32_t)));
    return MMC_OK;
}

template <int D>
int launch_cont(const MhContParams &p, bool replay, cudaStream_t stream) {
    const int block = 128;
    const unsigned grid = (unsigned)((p.chains + block - 1) / block);
    if (replay)
        mh_cont_kernel<D, true><<<grid, block, 0, stream>>>(p);
    else
        mh_cont_kernel<D, false><<<grid, block, 0, stream>>>(p);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

// IsotropicGaussian target of any dimension (<= kMhDynMax), f64 or f32 state: vectors in local memory, sums in the
// reference's sequential order
template <class T>
int launch_iso_dyn(const mmc_mh *h, const MhContParams &p, bool replay, cudaStream_t stream) {
    MMC_REQUIRE(h->target.kind == MMC_T_ISO_GAUSSIAN && h->dim <= kMhDynMax, "continuous MH: dim %d unsupported for this target (IsotropicGaussian runs up to %d)",
                h->dim, kMhDynMax);
    const T tstd = (T)h->target.params[0], pstd = (T)h->proposal.param;
    const IsoProposalF<T> q(h->proposal.param);
    const int block = 128;
    const unsigned grid = (unsigned)((p.chains + block - 1) / block);
    if (replay) mh_iso_dyn_kernel<T, true><<<grid, block, 0, stream>>>(p, tstd, pstd, q.ln_term);
    else mh_iso_dyn_kernel<T, false><<<grid, block, 0, stream>>>(p, tstd, pstd, q.ln_term);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

// MetropolisHastings<f32, f32, ..> (src/metropolis_hastings.rs:87): every operation in f32
int run_cont_f32(const mmc_mh *h, const MhContParams &p, bool replay, cudaStream_t stream) {
    const IsoProposalF<float> q(h->proposal.param);
    if (h->target.kind == MMC_T_GAUSSIAN2D)
        return launch_mh_functor<Gauss2DTargetF<float>, IsoProposalF<float>, float, 2>(Gauss2DTargetF<float>(h->target.params), q, p, replay, stream);
    switch (h->dim) {
    case 1: return launch_mh_functor<IsoTargetF<float, 1>, IsoProposalF<float>, float, 1>(IsoTargetF<float, 1>(h->target.params), q, p, replay, stream);
    case 2: return launch_mh_functor<IsoTargetF<float, 2>, IsoProposalF<float>, float, 2>(IsoTargetF<float, 2>(h->target.params), q, p, replay, stream);
    case 3: return launch_mh_functor<IsoTargetF<float, 3>, IsoProposalF<float>, float, 3>(IsoTargetF<float, 3>(h->target.params), q, p, replay, stream);
    case 4: return launch_mh_functor<IsoTargetF<float, 4>, IsoProposalF<float>, float, 4>(IsoTargetF<float, 4>(h->target.params), q, p, replay, stream);
    case 8: return launch_mh_functor<IsoTargetF<float, 8>, IsoProposalF<float>, float, 8>(IsoTargetF<float, 8>(h->target.params), q, p, replay, stream);
    default: return launch_iso_dyn<float>(h, p, replay, stream);
    }
}

int run_cont(mmc_mh *h, int64_t n_collect, int64_t n_discard, double *out_dev, const mmc_replay_mh *rp,
             cudaStream_t stream) {
    MhContParams p{};
    p.state = (double *)h->d_state;
    p.out = out_dev;
    p.noise = rp ? rp->noise : nullptr;
    p.u = rp ? rp->u : nullptr;
    p.trace = rp ? rp->trace : nullptr;
    p.chains = h->chains;
    p.chain_offset = h->chain_offset;
    p.step_base = h->step;
    p.n_collect = n_collect;
    p.n_discard = n_discard;
    p.out_pitch = h->out_pitch > 0 ? h->out_pitch : n_collect;
    p.key = seed_key(h->seed);
    p.target_kind = h->target.kind;
    for (int i = 0; i < 6; ++i) p.tp[i] = h->target.params[i];
    const double std_ = h->proposal.param;
    p.prop_std = std_;
    const double var = std_ * std_;
    p.prop_norm_term = -(double)h->dim * 0.5 * std::log(var * M_PI * std_ * std_);
    const bool replay = rp && rp->noise && rp->u;
    MMC_REQUIRE(!rp || replay, "MH replay needs both noise and u tapes");
    p.dim = h->dim;
    if (h->target.kind >= MMC_T_CUSTOM_BASE) {
        CustomTargetEntry e;
        MMC_REQUIRE(custom_target_get(h->target.kind, &e) && e.mh, "custom target kind %d is not registered for MH", h->target.kind);
        return e.mh(&p, replay ? 1 : 0, h->target.params, h->proposal.param, stream);
    }
    if (h->dtype == MMC_F32) return run_cont_f32(h, p, replay, stream);
    switch (h->dim) {
    case 1: return launch_cont<1>(p, replay, stream);
    case 2: return launch_cont<2>(p, replay, stream);
    case 3: return launch_cont<3>(p, replay, stream);
    case 4: return launch_cont<4>(p, replay, stream);
    case 8: return launch_cont<8>(p, replay, stream);
    default: return launch_iso_dyn<double>(h, p, replay, stream);
    }
}

// compact != 0: out_dev receives u8 (table_len <= 256) or u16 draws instead of u64 (host-API fast path)
int run_poisson(mmc_mh *h, int64_t n_collect, int64_t n_discard, void *out_dev, const mmc_replay_mh *rp,
                cudaStream_t stream, int64_t chain_begin = 0, int64_t chain_count = -1, bool compact = false) {
    if (chain_count < 0) chain_count = h->chains;
    MhPoissonParams p{};
    p.state = (uint64_t *)h->d_state + chain_begin;
    p.out = out_dev;
    p.flip = rp ? rp->flip : nullptr;
    p.u = rp ? rp->u : nullptr;
    p.lnfact = h->d_lnfact;
    p.lim = h->d_lim;
    p.table_len = h->table_len;
    p.lambda = h->target.params[0];
    p.ln_lambda = h->ln_lambda;
    p.ln_half = h->ln_half;
    p.chains = chain_count;
    p.chain_offset = h->chain_offset + chain_begin;
    p.step_base = h->step;
    p.n_collect = n_collect;
    p.n_discard = n_discard;
    p.out_pitch = (!compact && h->out_pitch > 0) ? h->out_pitch : n_collect;
    {
        uint32_t k0 = (uint32_t)h->seed, k1 = (uint32_t)(h->seed >> 32);
        for (int r = 0; r < 10; ++r) {
            p.rk[2 * r] = k0;
            p.rk[2 * r + 1] = k1;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
    }
    p.error_flag = h->d_error;
    p.force_up_at = h->proposal.kind == MMC_Q_REFLECT_RW ? 0xffffffffu : 0u;
    const bool replay = rp && rp->flip && rp->u;
    MMC_REQUIRE(!rp || replay, "Poisson MH replay needs both flip and u tapes");
    const int block = kPoisWarps * 32;
    const int64_t warps = (chain_count + 31) / 32;
    const unsigned grid = (unsigned)((warps + kPoisWarps - 1) / kPoisWarps);
    auto launch = [&](auto kernel, size_t tile_bytes) -> int {
        const size_t smem = (size_t)h->table_len * 8 + (size_t)kPoisWarps * tile_bytes;
        MMC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<grid, block, smem, stream>>>(p);
        MMC_CUDA(cudaGetLastError());
        return MMC_OK;
    };
    const bool thr = h->accept_mode != 0;
    const size_t tb64 = PoisTile<64, uint16_t>::kWarpBytes;
    if (replay) return thr ? launch(mh_poisson_kernel<true, true>, tb64) : launch(mh_poisson_kernel<true, false>, tb64);
    if (!thr) return launch(mh_poisson_kernel<false, false>, tb64);
    // native threshold path (the C2 hot loop): staging-tile shape selectable for tuning
    int tile = h->table_len <= 256 ? h->tile_u8 : h->tile_u16;
    if (const char *e = getenv("MMC_POIS_TILE")) tile = atoi(e);
    const bool u8 = h->table_len <= 256 && !getenv("MMC_POIS_U16");
    if (compact) {
        MMC_REQUIRE(!getenv("MMC_POIS_U16"), "compact path and MMC_POIS_U16 are exclusive");
        if (h->table_len <= 256)
            return launch(mh_poisson_kernel<false, true, 128, uint8_t, uint8_t>, PoisTile<128, uint8_t>::kWarpBytes);
        return launch(mh_poisson_kernel<false, true, 64, uint16_t, uint16_t>, PoisTile<64, uint16_t>::kWarpBytes);
    }
    if (u8) {
        if (tile >= 256) return launch(mh_poisson_kernel<false, true, 256, uint8_t>, PoisTile<256, uint8_t>::kWarpBytes);
        if (tile >= 128) return launch(mh_poisson_kernel<false, true, 128, uint8_t>, PoisTile<128, uint8_t>::kWarpBytes);
        return launch(mh_poisson_kernel<false, true, 64, uint8_t>, PoisTile<64, uint8_t>::kWarpBytes);
    }
    if (tile >= 256) return launch(mh_poisson_kernel<false, true, 256, uint16_t>, PoisTile<256, uint16_t>::kWarpBytes);
    if (tile >= 128) return launch(mh_poisson_kernel<false, true, 128, uint16_t>, PoisTile<128, uint16_t>::kWarpBytes);
    return launch(mh_poisson_kernel<false, true, 64, uint16_t>, tb64);
}

int check_error_flag(mmc_mh *h, cudaStream_t stream) {
    if (!h->d_error) return MMC_OK;
    int32_t flag = 0;
    MMC_CUDA(cudaMemcpyAsync(&flag, h->d_error, sizeof(flag), cudaMemcpyDeviceToHost, stream));
    MMC_CUDA(cudaStreamSynchronize(stream));
    if (flag) {
        set_error("Poisson MH: a chain left the tabulated state range [0, %d)", h->table_len - 1);
        return MMC_ERR_OVERFLOW;
    }
    return MMC_OK;
}

}  // namespace

extern "C" {

int mmc_mh_create(mmc_mh **out, const mmc_target_desc *target, const mmc_proposal_desc *proposal,
                  const void *init_host, int64_t chains, int32_t dim, int32_t state_dtype) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(out && target && proposal && init_host && chains > 0 && dim > 0, "mmc_mh_create: bad arguments");
    const bool poisson = target->kind == MMC_T_POISSON;
    MMC_REQUIRE(target->kind != MMC_T_CATEGORICAL, "use mmc_mh_create_categorical for the Categorical target");
    if (poisson) {
        MMC_REQUIRE(proposal->kind == MMC_Q_NONNEG_RW && state_dtype == MMC_U64 && dim == 1,
                    "Poisson target needs the nonnegative random-walk proposal, u64 state and dim 1");
    } else if (target->kind >= MMC_T_CUSTOM_BASE) {
        CustomTargetEntry e;
        const char *nm = "";
        MMC_REQUIRE(custom_target_get(target->kind, &e, &nm) && e.mh, "custom target kind %d is not registered for MH", target->kind);
        MMC_REQUIRE(e.dim == dim, "custom target '%s' has dim %d, got %d", nm, e.dim, dim);
        MMC_REQUIRE(state_dtype == MMC_F64, "custom MH targets run on f64 state");
    } else {
        MMC_REQUIRE(target->kind == MMC_T_GAUSSIAN2D || target->kind == MMC_T_ISO_GAUSSIAN,
                    "MH target kind %d is not a built-in MH target", target->kind);
        MMC_REQUIRE(proposal->kind == MMC_Q_ISO_GAUSSIAN && (state_dtype == MMC_F64 || state_dtype == MMC_F32),
                    "continuous MH needs the IsotropicGaussian proposal and f64 or f32 state");
        MMC_REQUIRE(target->kind != MMC_T_GAUSSIAN2D || dim == 2, "Gaussian2D needs dim 2");
        MMC_REQUIRE(dim <= kMhDynMax, "continuous MH runs up to dim %d, got %d", kMhDynMax, dim);
        MMC_REQUIRE(proposal->param > 0.0, "proposal std must be > 0");
    }
    mmc_mh *h = new mmc_mh();
    h->target = *target;
    h->proposal = *proposal;
    h->chains = chains;
    h->dim = dim;
    h->dtype = state_dtype;
    auto fail = [&](int code) { mmc_mh_destroy(h); return code; };
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__));
    const size_t bytes = (size_t)chains * dim * elem_size(state_dtype);
    e = cudaMalloc(&h->d_state, bytes);
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc(state)", __FILE__, __LINE__));
    e = cudaMemcpy(h->d_state, init_host, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMemcpy(state)", __FILE__, __LINE__));
    if (poisson) {
        rc = build_poisson_tables(h);
        if (rc) return fail(rc);
    }
    *out = h;
    return MMC_OK;
}

int mmc_mh_create_categorical(mmc_mh **out, const double *probs, int32_t n_categories, const void *init_host, int64_t chains) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(out && probs && init_host && chains > 0 && n_categories > 0 && n_categories <= 4000,
                "mmc_mh_create_categorical: bad arguments (1..4000 categories)");
    const uint64_t *init = static_cast<const uint64_t *>(init_host);
    for (int64_t c = 0; c < chains; ++c)
        MMC_REQUIRE(init[c] < (uint64_t)n_categories, "chain %lld starts at category %llu >= %d", (long long)c,
                    (unsigned long long)init[c], n_categories);
    mmc_mh *h = new mmc_mh();
    h->target.kind = MMC_T_CATEGORICAL;
    h->target.dim = 1;
    h->proposal.kind = MMC_Q_NONNEG_RW;
    h->chains = chains;
    h->dim = 1;
    h->dtype = MMC_U64;
    auto fail = [&](int code) { mmc_mh_destroy(h); return code; };
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__));
    e = cudaMalloc(&h->d_state, (size_t)chains * 8);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_state, init_host, (size_t)chains * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(cuda_fail(e, "mmc_mh_create_categorical", __FILE__, __LINE__));
    rc = build_categorical_tables(h, probs, n_categories);
    if (rc) return fail(rc);
    *out = h;
    return MMC_OK;
}

// Any Target<i32 / usize, f64> tabulated on [0, n_states) (-inf beyond) with one of the two +-1 random walks the
// reference's tests and examples use: NonnegativeProposal (examples/poisson_mh.rs:28-77) or the symmetric walk clamped to
// the support (PoissonRandomWalk / BinomialRandomWalk, tests/metrohast_poisson_test.rs:52-84,184-214).
int mmc_mh_create_tabulated(mmc_mh **out, const double *logp, int32_t n_states, int32_t proposal_kind, const void *init_host,
                            int64_t chains) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(out && logp && init_host && chains > 0 && n_states > 0 && n_states <= 4000,
                "mmc_mh_create_tabulated: bad arguments (1..4000 states)");
    MMC_REQUIRE(proposal_kind == MMC_Q_NONNEG_RW || proposal_kind == MMC_Q_REFLECT_RW,
                "mmc_mh_create_tabulated: proposal must be MMC_Q_NONNEG_RW or MMC_Q_REFLECT_RW");
    const uint64_t *init = static_cast<const uint64_t *>(init_host);
    for (int64_t c = 0; c < chains; ++c)
        MMC_REQUIRE(init[c] < (uint64_t)n_states, "chain %lld starts at state %llu >= %d", (long long)c,
                    (unsigned long long)init[c], n_states);
    mmc_mh *h = new mmc_mh();
    h->target.kind = MMC_T_TABULATED;
    h->target.dim = 1;
    h->proposal.kind = proposal_kind;
    h->chains = chains;
    h->dim = 1;
    h->dtype = MMC_U64;
    auto fail = [&](int code) { mmc_mh_destroy(h); return code; };
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__));
    e = cudaMalloc(&h->d_state, (size_t)chains * 8);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_state, init_host, (size_t)chains * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(cuda_fail(e, "mmc_mh_create_tabulated", __FILE__, __LINE__));
    const int64_t len = (int64_t)n_states + 16;   // 16 unreachable states past the support (see build_categorical_tables)
    h->table_len = (int32_t)len;
    std::vector<double> lnfact(len, 0.0), lp(len, -INFINITY);
    for (int32_t k = 0; k < n_states; ++k) lp[k] = logp[k];
    rc = upload_int_tables(h, lp, lnfact);
    if (rc) return fail(rc);
    *out = h;
    return MMC_OK;
}

int mmc_register_mh_target(const char *name, int32_t dim, mmc_mh_launch_fn fn) {
    MMC_REQUIRE(name && fn && dim > 0, "mmc_register_mh_target: bad arguments");
    return custom_target_register(name, dim, nullptr, nullptr, fn);
}

int mmc_mh_seed(mmc_mh *h, uint64_t seed) {
    MMC_REQUIRE(h, "null handle");
    h->seed = seed;
    h->step = 0;
    return MMC_OK;
}

int mmc_mh_set_chain_offset(mmc_mh *h, int64_t offset) {
    MMC_REQUIRE(h && offset >= 0, "bad chain offset");
    h->chain_offset = offset;
    return MMC_OK;
}

int mmc_mh_set_out_pitch(mmc_mh *h, int64_t pitch_steps) {
    MMC_REQUIRE(h && pitch_steps >= 0, "mmc_mh_set_out_pitch: bad arguments");
    h->out_pitch = pitch_steps;
    return MMC_OK;
}

int mmc_mh_set_accept_mode(mmc_mh *h, int32_t mode) {
    MMC_REQUIRE(h && (mode == 0 || mode == 1), "accept mode must be 0 or 1");
    MMC_REQUIRE(mode == 1 || h->target.kind == MMC_T_POISSON, "only the Poisson target has the device-evaluated accept mode");
    h->accept_mode = mode;
    return MMC_OK;
}

int mmc_mh_run_dev(mmc_mh *h, int64_t n_collect, int64_t n_discard, void *out_dev, const mmc_replay_mh *replay_dev,
                   void *stream) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0, "mmc_mh_run_dev: bad arguments");
    MMC_REQUIRE(out_dev || n_collect == 0, "mmc_mh_run_dev: out is null");
    MMC_REQUIRE(h->out_pitch == 0 || h->out_pitch >= n_collect, "mmc_mh_run_dev: out pitch %lld < n_collect", (long long)h->out_pitch);
    int rc;
    if (is_int_target(h))
        rc = run_poisson(h, n_collect, n_discard, (uint64_t *)out_dev, replay_dev, (cudaStream_t)stream);
    else
        rc = run_cont(h, n_collect, n_discard, (double *)out_dev, replay_dev, (cudaStream_t)stream);
    if (rc) return rc;
    h->step += n_collect + n_discard;
    return MMC_OK;
}

// Poisson host path: sample a block of chains into the compact (u8/u16) device buffer, copy it to pinned staging,
// and widen into the caller's u64 array on the host threads while the next block is sampled and copied.
static int mh_run_poisson_compact(mmc_mh *h, int64_t n_collect, int64_t n_discard, uint64_t *out_host) {
    const int eb = h->table_len <= 256 ? 1 : 2;
    const size_t row_bytes = (size_t)n_collect * eb;
    int64_t group = (int64_t)((size_t)(384u << 20) / (row_bytes ? row_bytes : 1));
    group = (group / 256) * 256;
    if (group < 256) group = 256;
    if (group > h->chains) group = h->chains;
    const size_t need = (size_t)group * row_bytes;
    if (h->stage_bytes < need) {
        for (int k = 0; k < 2; ++k) {
            if (h->d_compact[k]) cudaFree(h->d_compact[k]);
            if (h->h_stage[k]) cudaFreeHost(h->h_stage[k]);
            h->d_compact[k] = h->h_stage[k] = nullptr;
        }
        h->stage_bytes = 0;
        for (int k = 0; k < 2; ++k) {
            MMC_CUDA(cudaMalloc(&h->d_compact[k], need));
            MMC_CUDA(cudaHostAlloc(&h->h_stage[k], need, cudaHostAllocDefault));
            if (!h->pipe_stream[k]) MMC_CUDA(cudaStreamCreateWithFlags(&h->pipe_stream[k], cudaStreamNonBlocking));
            if (!h->pipe_event[k]) MMC_CUDA(cudaEventCreateWithFlags(&h->pipe_event[k], cudaEventDisableTiming));
        }
        h->stage_bytes = need;
    }
    const int64_t n_groups = (h->chains + group - 1) / group;
    auto widen = [&](int64_t g) {
        const int64_t begin = g * group, cnt = std::min(group, h->chains - begin);
        widen_to_u64(h->h_stage[g & 1], eb, out_host + begin * n_collect, (size_t)cnt * n_collect);
    };
    for (int64_t g = 0; g < n_groups; ++g) {
        const int k = (int)(g & 1);
        const int64_t begin = g * group, cnt = std::min(group, h->chains - begin);
        int rc = run_poisson(h, n_collect, n_discard, h->d_compact[k], nullptr, h->pipe_stream[k], begin, cnt, true);
        if (rc) return rc;
        MMC_CUDA(cudaMemcpyAsync(h->h_stage[k], h->d_compact[k], (size_t)cnt * row_bytes, cudaMemcpyDeviceToHost,
                                 h->pipe_stream[k]));
        MMC_CUDA(cudaEventRecord(h->pipe_event[k], h->pipe_stream[k]));
        if (g >= 1) {
            MMC_CUDA(cudaEventSynchronize(h->pipe_event[k ^ 1]));
            widen(g - 1);   // overlaps the sampling + copy of block g
        }
    }
    MMC_CUDA(cudaEventSynchronize(h->pipe_event[(n_groups - 1) & 1]));
    widen(n_groups - 1);
    h->step += n_collect + n_discard;
    return check_error_flag(h, h->pipe_stream[0]);
}

// Opt-in compact output for the integer targets: the draws reach the caller as the u8 / u16 values the kernel emits
// (1-2 B per draw over PCIe and in host memory instead of the reference API's 8 B `usize`), blocks of chains double
// buffered like mh_run_poisson_compact but copied straight into the caller's [chains, n_collect] array.
int mmc_mh_run_compact(mmc_mh *h, int64_t n_collect, int64_t n_discard, void *out_host, int32_t *elem_bytes) {
    MMC_REQUIRE(h && n_collect > 0 && n_discard >= 0 && out_host && elem_bytes, "mmc_mh_run_compact: bad arguments");
    MMC_REQUIRE(is_int_target(h) && h->accept_mode == 1, "mmc_mh_run_compact: integer targets with table accept mode only");
    MMC_REQUIRE(h->out_pitch == 0, "mmc_mh_run_compact: an output pitch is set on this handle");
    const int eb = h->table_len <= 256 ? 1 : 2;
    *elem_bytes = eb;
    const size_t row_bytes = (size_t)n_collect * eb;
    int64_t group = (int64_t)((size_t)(384u << 20) / row_bytes);
    group = (group / 256) * 256;
    if (group < 256) group = 256;
    if (group > h->chains) group = h->chains;
    const size_t need = (size_t)group * row_bytes;
    if (h->compact_only_bytes < need) {
        for (int k = 0; k < 2; ++k) {
            if (h->d_compact_only[k]) cudaFree(h->d_compact_only[k]);
            h->d_compact_only[k] = nullptr;
        }
        h->compact_only_bytes = 0;
        for (int k = 0; k < 2; ++k) {
            MMC_CUDA(cudaMalloc(&h->d_compact_only[k], need));
            if (!h->pipe_stream[k]) MMC_CUDA(cudaStreamCreateWithFlags(&h->pipe_stream[k], cudaStreamNonBlocking));
            if (!h->pipe_event[k]) MMC_CUDA(cudaEventCreateWithFlags(&h->pipe_event[k], cudaEventDisableTiming));
        }
        h->compact_only_bytes = need;
    }
    const int64_t n_groups = (h->chains + group - 1) / group;
    for (int64_t g = 0; g < n_groups; ++g) {
        const int k = (int)(g & 1);
        const int64_t begin = g * group, cnt = std::min(group, h->chains - begin);
        if (g >= 2) MMC_CUDA(cudaEventSynchronize(h->pipe_event[k]));   // the copy out of this device buffer has finished
        int rc = run_poisson(h, n_collect, n_discard, h->d_compact_only[k], nullptr, h->pipe_stream[k], begin, cnt, true);
        if (rc) return rc;
        MMC_CUDA(cudaMemcpyAsync(static_cast<unsigned char *>(out_host) + (size_t)begin * row_bytes, h->d_compact_only[k],
                                 (size_t)cnt * row_bytes, cudaMemcpyDeviceToHost, h->pipe_stream[k]));
        MMC_CUDA(cudaEventRecord(h->pipe_event[k], h->pipe_stream[k]));
    }
    MMC_CUDA(cudaStreamSynchronize(h->pipe_stream[0]));
    MMC_CUDA(cudaStreamSynchronize(h->pipe_stream[1]));
    h->step += n_collect + n_discard;
    return check_error_flag(h, h->pipe_stream[0]);
}

int mmc_mh_run(mmc_mh *h, int64_t n_collect, int64_t n_discard, void *out_host, const mmc_replay_mh *replay) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0 && (out_host || n_collect == 0), "mmc_mh_run: bad arguments");
    MMC_REQUIRE(h->out_pitch == 0, "mmc_mh_run: an output pitch only applies to mmc_mh_run_dev");
    MMC_CUDA(cudaDeviceSynchronize());  // earlier *_run_dev work on a caller stream may still be updating the chain state
    if (is_int_target(h) && !replay && h->accept_mode == 1 && n_collect > 0 && !getenv("MMC_NO_COMPACT"))
        return mh_run_poisson_compact(h, n_collect, n_discard, (uint64_t *)out_host);
    const int64_t steps = n_collect + n_discard;
    const size_t out_bytes = (size_t)h->chains * n_collect * h->dim * elem_size(h->dtype);
    int rc = grow(&h->d_out, &h->d_out_bytes, out_bytes ? out_bytes : 8);
    if (rc) return rc;
    mmc_replay_mh dev_rp{};
    const mmc_replay_mh *rp = nullptr;
    if (replay) {
        const bool poisson = is_int_target(h);
        const size_t n_u = (size_t)h->chains * steps;
        if (poisson) {
            MMC_REQUIRE(replay->flip && replay->u, "Poisson MH replay needs flip and u");
            if ((rc = grow(&h->d_replay[0], &h->d_replay_bytes[0], n_u ? n_u : 1))) return rc;
            MMC_CUDA(cudaMemcpyAsync(h->d_replay[0], replay->flip, n_u, cudaMemcpyHostToDevice, h->stream));
            dev_rp.flip = (const uint8_t *)h->d_replay[0];
        } else {
            MMC_REQUIRE(replay->noise && replay->u, "MH replay needs noise and u");
            const size_t nb = n_u * h->dim * 8;
            if ((rc = grow(&h->d_replay[0], &h->d_replay_bytes[0], nb ? nb : 8))) return rc;
            MMC_CUDA(cudaMemcpyAsync(h->d_replay[0], replay->noise, nb, cudaMemcpyHostToDevice, h->stream));
            dev_rp.noise = (const double *)h->d_replay[0];
            if (replay->trace) {
                if ((rc = grow((void **)&h->d_trace, &h->d_trace_bytes, n_u * 4 * 8))) return rc;
                dev_rp.trace = h->d_trace;
            }
        }
        if ((rc = grow(&h->d_replay[1], &h->d_replay_bytes[1], n_u ? n_u * 8 : 8))) return rc;
        MMC_CUDA(cudaMemcpyAsync(h->d_replay[1], replay->u, n_u * 8, cudaMemcpyHostToDevice, h->stream));
        dev_rp.u = (const double *)h->d_replay[1];
        rp = &dev_rp;
    }
    rc = mmc_mh_run_dev(h, n_collect, n_discard, h->d_out, rp, h->stream);
    if (rc) return rc;
    if (out_bytes) MMC_CUDA(cudaMemcpyAsync(out_host, h->d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (rp && rp->trace)
        MMC_CUDA(cudaMemcpyAsync(replay->trace, h->d_trace, (size_t)h->chains * steps * 4 * 8, cudaMemcpyDeviceToHost,
                                 h->stream));
    MMC_CUDA(cudaStreamSynchronize(h->stream));
    return check_error_flag(h, h->stream);
}

int mmc_mh_d2h_bytes_per_draw(mmc_mh *h) {
    if (!h) return MMC_ERR_INVALID;
    if (is_int_target(h) && h->accept_mode == 1 && !getenv("MMC_NO_COMPACT")) return h->table_len <= 256 ? 1 : 2;
    return (int)elem_size(h->dtype) * h->dim;
}

int mmc_mh_run_progress(mmc_mh *h, int64_t n_collect, int64_t n_discard, void *out_host, int64_t block, mmc_progress_fn cb,
                        void *user, mmc_run_stats *stats) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0 && (out_host || n_collect == 0), "mmc_mh_run_progress: bad arguments");
    MMC_REQUIRE(h->out_pitch == 0, "mmc_mh_run_progress: an output pitch is set on this handle");
    ProgressSpec sp{h->chains, h->dim, h->dtype, MMC_TRACK_PER_CHAIN, true, h->d_state};
    auto run_block = [&](int64_t k, void *dst, int64_t pitch, bool) {
        h->out_pitch = pitch;
        const int rc = mmc_mh_run_dev(h, k, 0, dst, nullptr, h->stream);
        h->out_pitch = 0;
        return rc;
    };
    auto discard = [&](int64_t k) { return mmc_mh_run_dev(h, 0, k, nullptr, nullptr, h->stream); };
    return run_progress_blocks(sp, n_collect, n_discard, out_host, block, cb, user, stats, h->stream, run_block, discard);
}

int mmc_mh_get_state(mmc_mh *h, void *state_host) {
    MMC_REQUIRE(h && state_host, "mmc_mh_get_state: bad arguments");
    MMC_CUDA(cudaDeviceSynchronize());  // orders the copy after *_run_dev work on any caller stream
    MMC_CUDA(cudaMemcpy(state_host, h->d_state, (size_t)h->chains * h->dim * elem_size(h->dtype), cudaMemcpyDeviceToHost));
    return check_error_flag(h, h->stream);
}

int mmc_mh_set_state(mmc_mh *h, const void *state_host) {
    MMC_REQUIRE(h && state_host, "mmc_mh_set_state: bad arguments");
    MMC_CUDA(cudaDeviceSynchronize());
    MMC_CUDA(cudaMemcpy(h->d_state, state_host, (size_t)h->chains * h->dim * elem_size(h->dtype), cudaMemcpyHostToDevice));
    if (h->d_error) MMC_CUDA(cudaMemset(h->d_error, 0, sizeof(int32_t)));
    return MMC_OK;
}

void mmc_mh_destroy(mmc_mh *h) {
    if (!h) return;
    cudaFree(h->d_state);
    cudaFree(h->d_lnfact);
    cudaFree(h->d_lim);
    cudaFree(h->d_error);
    cudaFree(h->d_out);
    for (auto p : h->d_replay) cudaFree(p);
    cudaFree(h->d_trace);
    for (int k = 0; k < 2; ++k) {
        cudaFree(h->d_compact[k]);
        cudaFree(h->d_compact_only[k]);
        if (h->h_stage[k]) cudaFreeHost(h->h_stage[k]);
        if (h->pipe_stream[k]) cudaStreamDestroy(h->pipe_stream[k]);
        if (h->pipe_event[k]) cudaEventDestroy(h->pipe_event[k]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

}  // extern "C"
