// HMC C ABI (see include/minimcmc.h) — host side of K2 and the built-in target dispatch.
#include <cmath>
#include <mutex>
#include <string>
#include <vector>

#include "mmc_dense.cuh"
#include <type_traits>

#include "mmc_hmc.cuh"
#include "mmc_hmc_pair.cuh"
#include "mmc_progress.cuh"
#include "mmc_hmc_warp.cuh"
#include "mmc_targets.cuh"

using namespace mmc;

struct mmc_hmc {
    mmc_target_desc target{};
    int64_t chains = 0;
    int32_t dim = 0;
    double step_size = 0;
    int32_t n_leapfrog = 0;
    int64_t chain_offset = 0;
    int64_t step = 0;
    uint64_t seed = 0;
    int32_t exact = 0;
    int64_t out_pitch = 0;       // *_dev runs: draws per chain row of the caller's tensor (0 = n_collect)
    int32_t gemm_path = -1;      // dense Gaussian: -1 = auto (tcgen05 CTA pairs, mixed split), 0 = FP32 SIMT tiles, 1 / 2 = tcgen05 3xTF32, 3 = TF32 + BF16 mixed split
    DenseState *dense = nullptr;  // only for MMC_T_DENSE_GAUSSIAN
    float *d_pos = nullptr;
    unsigned long long *d_accept = nullptr;
    int64_t total_transitions = 0;
    cudaStream_t stream = nullptr;
    float *d_out = nullptr;
    size_t d_out_bytes = 0;
    float *d_mom = nullptr, *d_u = nullptr, *d_trace = nullptr;
    size_t d_mom_bytes = 0, d_u_bytes = 0, d_trace_bytes = 0;
};

namespace {

int grow(float **ptr, size_t *cap, size_t need) {
    if (*cap >= need) return MMC_OK;
    if (*ptr) MMC_CUDA(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    MMC_CUDA(cudaMalloc((void **)ptr, need));
    *cap = need;
    return MMC_OK;
}

// DiffableGaussian2D::new, src/distributions.rs:227-251 (T = f64), cast to the backend float afterwards
template <class A>
DiffGaussian2D<A> make_diff_gaussian(const mmc_target_desc &t) {
    const double c00 = t.params[2], c01 = t.params[3], c10 = t.params[4], c11 = t.params[5];
    const double det = c00 * c11 - c01 * c10;
    const double inv_det = 1.0 / det;
    DiffGaussian2D<A> g;
    g.m0 = (float)t.params[0];
    g.m1 = (float)t.params[1];
    g.p00 = (float)(c11 * inv_det);
    g.p01 = (float)(-c01 * inv_det);
    g.p10 = (float)(-c10 * inv_det);
    g.p11 = (float)(c00 * inv_det);
    const double two = 2.0;
    g.norm_const = (float)(-(two * std::log(two * M_PI) + std::log(det)) / two);
    return g;
}

template <class A>
int dispatch(const mmc_hmc *h, const HmcParams &p, bool replay, cudaStream_t s) {
    const mmc_target_desc &t = h->target;
    switch (t.kind) {
    case MMC_T_ROSENBROCK_ND:
        // throughput mode: two chains per thread on packed f32x2 instructions (mmc_hmc_pair.cuh), for native draws
        // AND for replay / trace runs, so that the parity tests exercise the kernel that produces the throughput
        // numbers; the scalar kernel serves exact runs (and MMC_HMC_NO_PAIR=1 for A/B comparisons)
        if (std::is_same<A, Fast>::value && !getenv("MMC_HMC_NO_PAIR")) {
            switch (t.dim) {
            case 2: return launch_hmc_pair<RosenbrockND2<2>>({}, p, replay, s);
            case 3: return launch_hmc_pair<RosenbrockND2<3>>({}, p, replay, s);
            case 4: return launch_hmc_pair<RosenbrockND2<4>>({}, p, replay, s);
            case 5: return launch_hmc_pair<RosenbrockND2<5>>({}, p, replay, s);
            case 8: return launch_hmc_pair<RosenbrockND2<8>>({}, p, replay, s);
            case 10: return launch_hmc_pair<RosenbrockND2<10>>({}, p, replay, s);
            case 16: return launch_hmc_pair<RosenbrockND2<16>>({}, p, replay, s);
            default: break;
            }
        }
        switch (t.dim) {
        case 2: return launch_hmc<RosenbrockND<A, 2>, A>({}, p, replay, s);
        case 3: return launch_hmc<RosenbrockND<A, 3>, A>({}, p, replay, s);
        case 4: return launch_hmc<RosenbrockND<A, 4>, A>({}, p, replay, s);
        case 5: return launch_hmc<RosenbrockND<A, 5>, A>({}, p, replay, s);
        case 8: return launch_hmc<RosenbrockND<A, 8>, A>({}, p, replay, s);
        case 10: return launch_hmc<RosenbrockND<A, 10>, A>({}, p, replay, s);
        case 16: return launch_hmc<RosenbrockND<A, 16>, A>({}, p, replay, s);
        default: break;
        }
        break;
    case MMC_T_ROSENBROCK_2D: {
        Rosenbrock2D<A> r;
        r.a = (float)t.params[0];
        r.b = (float)t.params[1];
        return launch_hmc<Rosenbrock2D<A>, A>(r, p, replay, s);
    }
    case MMC_T_DIFF_GAUSSIAN2D: return launch_hmc<DiffGaussian2D<A>, A>(make_diff_gaussian<A>(t), p, replay, s);
    case MMC_T_STD_NORMAL:
        switch (t.dim) {
        case 1: return launch_hmc<StdNormal<A, 1>, A>({}, p, replay, s);
        case 2: return launch_hmc<StdNormal<A, 2>, A>({}, p, replay, s);
        case 3: return launch_hmc<StdNormal<A, 3>, A>({}, p, replay, s);
        case 4: return launch_hmc<StdNormal<A, 4>, A>({}, p, replay, s);
        default: break;
        }
        break;
    default: break;
    }
    // other dimensions: one chain per warp (D <= 512)
    const int D = t.dim;
    if (t.kind == MMC_T_ROSENBROCK_ND) {
        if (D <= 32) return launch_hmc_warp<WRosenbrockND<A, 1>, A, 1>({D}, p, D, replay, s);
        if (D <= 128) return launch_hmc_warp<WRosenbrockND<A, 4>, A, 4>({D}, p, D, replay, s);
        if (D <= 256) return launch_hmc_warp<WRosenbrockND<A, 8>, A, 8>({D}, p, D, replay, s);
        if (D <= 512) return launch_hmc_warp<WRosenbrockND<A, 16>, A, 16>({D}, p, D, replay, s);
    } else if (t.kind == MMC_T_STD_NORMAL) {
        if (D <= 32) return launch_hmc_warp<WStdNormal<A, 1>, A, 1>({D}, p, D, replay, s);
        if (D <= 128) return launch_hmc_warp<WStdNormal<A, 4>, A, 4>({D}, p, D, replay, s);
        if (D <= 512) return launch_hmc_warp<WStdNormal<A, 16>, A, 16>({D}, p, D, replay, s);
    }
    set_error("HMC: target kind %d with dim %d is not compiled in (register kernel: listed dims; warp kernel: dim <= 512)",
              t.kind, t.dim);
    return MMC_ERR_UNSUPPORTED;
}

template <int D>
int export_tape(const mmc_hmc *h, int64_t step_base, int64_t steps, float *mom, float *u, cudaStream_t s) {
    const int64_t total = h->chains * steps;
    const int block = 256;
    hmc_export_tape_kernel<D><<<(unsigned)((total + block - 1) / block), block, 0, s>>>(
        seed_key(h->seed), h->chains, h->chain_offset, step_base, steps, mom, u);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

}  // namespace

extern "C" {

int mmc_hmc_create(mmc_hmc **out, const mmc_target_desc *target, const float *init_host, int64_t chains, int32_t dim,
                   double step_size, int32_t n_leapfrog) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(out && target && init_host && chains > 0 && dim > 0 && n_leapfrog >= 0,
                "mmc_hmc_create: bad arguments");
    MMC_REQUIRE(target->dim == dim, "target dim %d != dim %d", target->dim, dim);
    if (target->kind >= MMC_T_CUSTOM_BASE) {
        CustomTargetEntry e;
        const char *nm = "";
        MMC_REQUIRE(custom_target_get(target->kind, &e, &nm) && e.hmc, "custom target kind %d is not registered for HMC", target->kind);
        MMC_REQUIRE(e.dim == dim, "custom target '%s' has dim %d, got %d", nm, e.dim, dim);
    }
    mmc_hmc *h = new mmc_hmc();
    h->target = *target;
    h->chains = chains;
    h->dim = dim;
    h->step_size = step_size;
    h->n_leapfrog = n_leapfrog;
    auto fail = [&](cudaError_t e, const char *what) {
        int code = cuda_fail(e, what, __FILE__, __LINE__);
        mmc_hmc_destroy(h);
        return code;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    const size_t bytes = (size_t)chains * dim * sizeof(float);
    if ((e = cudaMalloc((void **)&h->d_pos, bytes)) != cudaSuccess) return fail(e, "cudaMalloc(positions)");
    if ((e = cudaMemcpy(h->d_pos, init_host, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "memcpy");
    if ((e = cudaMalloc((void **)&h->d_accept, 8)) != cudaSuccess) return fail(e, "cudaMalloc(counter)");
    if ((e = cudaMemset(h->d_accept, 0, 8)) != cudaSuccess) return fail(e, "memset");
    if (target->kind == MMC_T_DENSE_GAUSSIAN) {
        rc = dense_create(&h->dense, target, chains);
        if (rc) { mmc_hmc_destroy(h); return rc; }
        h->target.vec = nullptr;  // host pointers are not kept
        h->target.mat = nullptr;
    }
    *out = h;
    return MMC_OK;
}

int mmc_register_hmc_target(const char *name, int32_t dim, mmc_hmc_launch_fn fn) {
    MMC_REQUIRE(name && fn && dim > 0, "mmc_register_hmc_target: bad arguments");
    return custom_target_register(name, dim, fn, nullptr, nullptr);
}

int mmc_lookup_target(const char *name) { return custom_target_lookup(name); }

int mmc_hmc_set_seed(mmc_hmc *h, uint64_t seed) {
    MMC_REQUIRE(h, "null handle");
    h->seed = seed;
    h->step = 0;
    return MMC_OK;
}

int mmc_hmc_set_chain_offset(mmc_hmc *h, int64_t offset) {
    MMC_REQUIRE(h && offset >= 0, "bad chain offset");
    h->chain_offset = offset;
    return MMC_OK;
}

int mmc_hmc_set_exact(mmc_hmc *h, int32_t exact) {
    MMC_REQUIRE(h, "null handle");
    h->exact = exact ? 1 : 0;
    return MMC_OK;
}

int mmc_hmc_set_out_pitch(mmc_hmc *h, int64_t pitch_steps) {
    MMC_REQUIRE(h && pitch_steps >= 0, "mmc_hmc_set_out_pitch: bad arguments");
    h->out_pitch = pitch_steps;
    return MMC_OK;
}

int mmc_hmc_set_gemm_path(mmc_hmc *h, int32_t path) {
    MMC_REQUIRE(h && path >= -1 && path <= 3, "gemm path must be -1 (auto), 0 (FP32 SIMT), 1 (tcgen05 3xTF32), 2 (tcgen05 3xTF32, CTA pairs) or 3 (CTA pairs, TF32 + BF16 mixed split)");
    h->gemm_path = path;
    return MMC_OK;
}

int mmc_hmc_run_dev(mmc_hmc *h, int64_t n_collect, int64_t n_discard, float *out_dev, const mmc_replay_hmc *rp,
                    void *stream) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0, "mmc_hmc_run_dev: bad arguments");
    MMC_REQUIRE(h->out_pitch == 0 || h->out_pitch >= n_collect, "mmc_hmc_run_dev: out pitch %lld < n_collect", (long long)h->out_pitch);
    const bool replay = rp && rp->momenta && rp->u;
    MMC_REQUIRE(!rp || replay, "HMC replay needs both momenta and u tapes");
    if (h->dense) {
        DenseRunArgs a{};
        a.positions = h->d_pos;
        a.out = n_collect > 0 ? out_dev : nullptr;
        a.momenta = replay ? rp->momenta : nullptr;
        a.u = replay ? rp->u : nullptr;
        a.trace = rp ? rp->trace : nullptr;
        a.accept_count = h->d_accept;
        a.chains = h->chains;
        a.chain_offset = h->chain_offset;
        a.step_base = h->step;
        a.n_collect = n_collect;
        a.n_discard = n_discard;
        a.out_pitch = h->out_pitch > 0 ? h->out_pitch : n_collect;
        a.eps = (float)h->step_size;
        a.n_leapfrog = h->n_leapfrog;
        a.seed = h->seed;
        // the dense contraction goes through the tensor cores by default (north_star); exact runs and dimensions the
        // 128 x 256 tcgen05 tiles do not divide use the FP32 SIMT tiles
        // auto: tcgen05 CTA pairs (any dim: the rows are zero padded to whole 256-column tiles); dim 128 and the reference
        // arithmetic run on the FP32 SIMT tiles
        a.gemm_path = (h->exact || h->dim == 128) ? 0 : (h->gemm_path >= 0 ? h->gemm_path : 3);
        int rc = dense_run(h->dense, a, (cudaStream_t)stream);
        if (rc) return rc;
        h->step += n_collect + n_discard;
        h->total_transitions += (n_collect + n_discard) * h->chains;
        return MMC_OK;
    }
    HmcParams p{};
    p.positions = h->d_pos;
    p.out = n_collect > 0 ? out_dev : nullptr;
    p.momenta = replay ? rp->momenta : nullptr;
    p.u = replay ? rp->u : nullptr;
    p.trace = rp ? rp->trace : nullptr;
    p.accept_count = h->d_accept;
    p.chains = h->chains;
    p.chain_offset = h->chain_offset;
    p.step_base = h->step;
    p.n_collect = n_collect;
    p.n_discard = n_discard;
    p.out_pitch = h->out_pitch > 0 ? h->out_pitch : n_collect;
    p.eps = (float)h->step_size;
    p.n_leapfrog = h->n_leapfrog;
    p.key = seed_key(h->seed);
    int rc;
    if (h->target.kind >= MMC_T_CUSTOM_BASE) {
        CustomTargetEntry e;
        mmc_hmc_launch_fn fn = custom_target_get(h->target.kind, &e) ? e.hmc : nullptr;
        MMC_REQUIRE(fn, "custom target kind %d is not registered", h->target.kind);
        rc = fn(&p, replay ? 1 : 0, h->exact, h->target.params, stream);
    } else {
        rc = h->exact ? dispatch<Exact>(h, p, replay, (cudaStream_t)stream)
                      : dispatch<Fast>(h, p, replay, (cudaStream_t)stream);
    }
    if (rc) return rc;
    h->step += n_collect + n_discard;
    h->total_transitions += (n_collect + n_discard) * h->chains;
    return MMC_OK;
}

int mmc_hmc_step(mmc_hmc *h) {
    MMC_REQUIRE(h, "null handle");
    int rc = mmc_hmc_run_dev(h, 0, 1, nullptr, nullptr, h->stream);
    if (rc) return rc;
    MMC_CUDA(cudaStreamSynchronize(h->stream));
    return MMC_OK;
}

int mmc_hmc_run(mmc_hmc *h, int64_t n_collect, int64_t n_discard, float *out_host, const mmc_replay_hmc *replay) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0 && (out_host || n_collect == 0), "mmc_hmc_run: bad arguments");
    MMC_REQUIRE(h->out_pitch == 0, "mmc_hmc_run: an output pitch only applies to mmc_hmc_run_dev");
    const int64_t steps = n_collect + n_discard;
    const size_t out_bytes = (size_t)h->chains * n_collect * h->dim * sizeof(float);
    int rc = grow(&h->d_out, &h->d_out_bytes, out_bytes ? out_bytes : 16);
    if (rc) return rc;
    mmc_replay_hmc dev_rp{};
    const mmc_replay_hmc *rp = nullptr;
    if (replay) {
        MMC_REQUIRE(replay->momenta && replay->u, "HMC replay needs momenta and u");
        const size_t nu = (size_t)steps * h->chains;
        if ((rc = grow(&h->d_mom, &h->d_mom_bytes, nu * h->dim * 4 + 16))) return rc;
        if ((rc = grow(&h->d_u, &h->d_u_bytes, nu * 4 + 16))) return rc;
        MMC_CUDA(cudaMemcpyAsync(h->d_mom, replay->momenta, nu * h->dim * 4, cudaMemcpyHostToDevice, h->stream));
        MMC_CUDA(cudaMemcpyAsync(h->d_u, replay->u, nu * 4, cudaMemcpyHostToDevice, h->stream));
        dev_rp.momenta = h->d_mom;
        dev_rp.u = h->d_u;
        if (replay->trace) {
            if ((rc = grow(&h->d_trace, &h->d_trace_bytes, nu * 16 + 16))) return rc;
            dev_rp.trace = h->d_trace;
        }
        rp = &dev_rp;
    }
    rc = mmc_hmc_run_dev(h, n_collect, n_discard, h->d_out, rp, h->stream);
    if (rc) return rc;
    if (out_bytes) MMC_CUDA(cudaMemcpyAsync(out_host, h->d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (rp && rp->trace)
        MMC_CUDA(cudaMemcpyAsync(replay->trace, h->d_trace, (size_t)steps * h->chains * 16, cudaMemcpyDeviceToHost,
                                 h->stream));
    MMC_CUDA(cudaStreamSynchronize(h->stream));
    return MMC_OK;
}

int mmc_hmc_run_progress(mmc_hmc *h, int64_t n_collect, int64_t n_discard, float *out_host, int64_t block, mmc_progress_fn cb,
                         void *user, mmc_run_stats *stats) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0 && (out_host || n_collect == 0), "mmc_hmc_run_progress: bad arguments");
    MMC_REQUIRE(h->out_pitch == 0, "mmc_hmc_run_progress: an output pitch is set on this handle");
    ProgressSpec sp{h->chains, h->dim, MMC_F32, MMC_TRACK_MULTI, false, h->d_pos};
    auto run_block = [&](int64_t k, void *dst, int64_t pitch, bool) {
        h->out_pitch = pitch;
        const int rc = mmc_hmc_run_dev(h, k, 0, static_cast<float *>(dst), nullptr, h->stream);
        h->out_pitch = 0;
        return rc;
    };
    auto discard = [&](int64_t k) { return mmc_hmc_run_dev(h, 0, k, nullptr, nullptr, h->stream); };
    return run_progress_blocks(sp, n_collect, n_discard, out_host, block, cb, user, stats, h->stream, run_block, discard);
}

int mmc_hmc_get_positions(mmc_hmc *h, float *positions_host) {
    MMC_REQUIRE(h && positions_host, "mmc_hmc_get_positions: bad arguments");
    MMC_CUDA(cudaStreamSynchronize(h->stream));
    MMC_CUDA(cudaMemcpy(positions_host, h->d_pos, (size_t)h->chains * h->dim * sizeof(float), cudaMemcpyDeviceToHost));
    return MMC_OK;
}

int mmc_hmc_positions_dev(mmc_hmc *h, float **positions_dev) {
    MMC_REQUIRE(h && positions_dev, "mmc_hmc_positions_dev: bad arguments");
    *positions_dev = h->d_pos;
    return MMC_OK;
}

int mmc_hmc_get_accept_counts(mmc_hmc *h, int64_t *accepted, int64_t *total) {
    MMC_REQUIRE(h, "null handle");
    unsigned long long a = 0;
    MMC_CUDA(cudaDeviceSynchronize());
    MMC_CUDA(cudaMemcpy(&a, h->d_accept, 8, cudaMemcpyDeviceToHost));
    if (accepted) *accepted = (int64_t)a;
    if (total) *total = h->total_transitions;
    return MMC_OK;
}

int mmc_hmc_export_tape_dev(mmc_hmc *h, int64_t step_base, int64_t steps, float *momenta_dev, float *u_dev,
                            void *stream) {
    MMC_REQUIRE(h && momenta_dev && u_dev && steps > 0, "mmc_hmc_export_tape_dev: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (h->dense)
        return dense_export_tape(h->chains, h->dim, h->chain_offset, h->seed, step_base, steps, momenta_dev, u_dev, s);
    switch (h->dim) {
    case 1: return export_tape<1>(h, step_base, steps, momenta_dev, u_dev, s);
    case 2: return export_tape<2>(h, step_base, steps, momenta_dev, u_dev, s);
    case 3: return export_tape<3>(h, step_base, steps, momenta_dev, u_dev, s);
    case 4: return export_tape<4>(h, step_base, steps, momenta_dev, u_dev, s);
    case 5: return export_tape<5>(h, step_base, steps, momenta_dev, u_dev, s);
    case 8: return export_tape<8>(h, step_base, steps, momenta_dev, u_dev, s);
    case 10: return export_tape<10>(h, step_base, steps, momenta_dev, u_dev, s);
    case 16: return export_tape<16>(h, step_base, steps, momenta_dev, u_dev, s);
    default: {
        const int64_t total = steps * h->chains * ((h->dim + 3) / 4);
        hmc_export_tape_any_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(seed_key(h->seed), h->chains, h->dim,
                                                                                   h->chain_offset, step_base, steps,
                                                                                   momenta_dev, u_dev);
        MMC_CUDA(cudaGetLastError());
        return MMC_OK;
    }
    }
}

void mmc_hmc_destroy(mmc_hmc *h) {
    if (!h) return;
    dense_destroy(h->dense);
    cudaFree(h->d_pos);
    cudaFree(h->d_accept);
    cudaFree(h->d_out);
    cudaFree(h->d_mom);
    cudaFree(h->d_u);
    cudaFree(h->d_trace);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

}  // extern "C"
