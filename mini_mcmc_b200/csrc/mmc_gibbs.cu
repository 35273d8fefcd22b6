// Gibbs sampler (src/gibbs.rs:89-205): GibbsMarkovChain::step sweeps the coordinates in order,
// state[i] = target.sample(i, &state); GibbsSampler runs independent chains through ChainRunner::run.
// One thread per chain for the whole run, f64 state in registers, draws written to [chains, n_collect, dim].
// Built-in conditionals: ConstantConditional (src/gibbs.rs:218-226) and the two-component Gaussian mixture of the
// reference's tests and examples/mixture_gibbs.rs:24-72 (state = [x, z]).  Compiled with -fmad=false: the f64
// operation sequence is the reference's.  RNG contract (native mode): Philox counter (chain, step, sub = coordinate):
// words (0,1) -> 53-bit uniform, Box-Muller on words (0,1),(2,3) -> normal z-score.
#include <mutex>
#include <string>
#include <vector>

#include "mmc_gibbs.cuh"
#include "mmc_progress.cuh"

namespace mmc {
namespace {


__device__ __forceinline__ double mix_pdf(double x, double mu, double sigma) {
    const double var = sigma * sigma;
    const double coeff = 1.0 / sqrt(2.0 * 3.14159265358979323846 * var);
    const double d = x - mu;
    const double exp_val = exp(-(d * d) / (2.0 * var));
    return coeff * exp_val;
}

template <bool kReplay>
__global__ void __launch_bounds__(128) gibbs_mixture_kernel(const GibbsParams p) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.chains) return;
    double x = p.state[c * 2], z = p.state[c * 2 + 1];
    const int64_t steps = p.n_collect + p.n_discard;
    const uint64_t gc = (uint64_t)(c + p.chain_offset);
    for (int64_t s = 0; s < steps; ++s) {
        double zs, u;
        if (kReplay) {
            zs = p.normals[c * steps + s];
            u = p.unifs[c * steps + s];
        } else {
            const uint32_t step = (uint32_t)(p.step_base + s);
            const uint4 w0 = philox4x32_10(p.key, make_uint4((uint32_t)gc, (uint32_t)(gc >> 32), step, 0u));
            double n1;
            box_muller_f64(w0, zs, n1);
            const uint4 w1 = philox4x32_10(p.key, make_uint4((uint32_t)gc, (uint32_t)(gc >> 32), step, 1u));
            u = u53_half_open(w1.x, w1.y);
        }
        // i = 0: x | z ~ Normal(mu_z, sigma_z) = mean + std * zscore
        x = (z < 0.5) ? p.p[0] + p.p[1] * zs : p.p[2] + p.p[3] * zs;
        // i = 1: z | x
        const double p0 = p.p[4] * mix_pdf(x, p.p[0], p.p[1]);
        const double p1 = (1.0 - p.p[4]) * mix_pdf(x, p.p[2], p.p[3]);
        const double total = p0 + p1;
        const double prob_z1 = total > 0.0 ? p1 / total : 0.5;
        z = (u < prob_z1) ? 1.0 : 0.0;
        if (p.trace) {
            p.trace[(c * steps + s) * 2] = zs;
            p.trace[(c * steps + s) * 2 + 1] = u;
        }
        if (s >= p.n_discard) {
            double2 *o = reinterpret_cast<double2 *>(p.out + (c * p.out_pitch + (s - p.n_discard)) * 2);
            *o = make_double2(x, z);
        }
    }
    p.state[c * 2] = x;
    p.state[c * 2 + 1] = z;
}

__global__ void __launch_bounds__(128) gibbs_constant_kernel(const GibbsParams p) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.chains) return;
    const int64_t steps = p.n_collect + p.n_discard;
    if (steps == 0) return;
    for (int i = 0; i < p.D; ++i) p.state[c * p.D + i] = p.p[0];
    for (int64_t s = p.n_discard; s < steps; ++s)
        for (int i = 0; i < p.D; ++i) p.out[(c * p.out_pitch + (s - p.n_discard)) * p.D + i] = p.p[0];
}

}  // namespace
}  // namespace mmc

using namespace mmc;

struct mmc_gibbs {
    mmc_conditional_desc cond{};
    int64_t chains = 0;
    int32_t dim = 0;
    int64_t chain_offset = 0, step = 0, out_pitch = 0;
    uint64_t seed = 0;
    double *d_state = nullptr;
    cudaStream_t stream = nullptr;
    double *d_out = nullptr;
    size_t d_out_bytes = 0;
    double *d_tape[3] = {nullptr, nullptr, nullptr};
    size_t d_tape_bytes[3] = {0, 0, 0};
};

namespace {
struct CustomConditional {
    std::string name;
    int dim;
    mmc_gibbs_launch_fn fn;
};
std::mutex g_cond_mutex;
std::vector<CustomConditional> &cond_registry() {
    static std::vector<CustomConditional> r;
    return r;
}

int grow(double **buf, size_t *have, size_t want) {
    if (want <= *have) return MMC_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
    MMC_CUDA(cudaMalloc((void **)buf, want));
    *have = want;
    return MMC_OK;
}
}  // namespace

extern "C" {

void mmc_gibbs_destroy(mmc_gibbs *h) {
    if (!h) return;
    cudaFree(h->d_state);
    cudaFree(h->d_out);
    for (auto *t : h->d_tape) cudaFree(t);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int mmc_gibbs_create(mmc_gibbs **out, const mmc_conditional_desc *cond, const double *init_host, int64_t chains, int32_t dim) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(out && cond && init_host && chains > 0 && dim > 0, "mmc_gibbs_create: bad arguments");
    if (cond->kind >= MMC_G_CUSTOM_BASE) {
        std::lock_guard<std::mutex> lock(g_cond_mutex);
        const size_t idx = (size_t)(cond->kind - MMC_G_CUSTOM_BASE);
        MMC_REQUIRE(idx < cond_registry().size(), "conditional kind %d is not registered", cond->kind);
        MMC_REQUIRE(cond_registry()[idx].dim == dim, "conditional '%s' has dim %d, the initial states have dim %d",
                    cond_registry()[idx].name.c_str(), cond_registry()[idx].dim, dim);
    } else {
        MMC_REQUIRE(cond->kind == MMC_G_CONSTANT || cond->kind == MMC_G_MIXTURE2, "unknown conditional kind %d", cond->kind);
    }
    MMC_REQUIRE(cond->kind != MMC_G_MIXTURE2 || dim == 2, "the mixture conditional has state [x, z]: dim must be 2");
    MMC_REQUIRE(cond->kind != MMC_G_MIXTURE2 || (cond->params[1] > 0.0 && cond->params[3] > 0.0), "mixture std deviations must be > 0");
    mmc_gibbs *h = new mmc_gibbs();
    h->cond = *cond;
    h->chains = chains;
    h->dim = dim;
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    const size_t bytes = (size_t)chains * dim * sizeof(double);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_state, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_state, init_host, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        int code = cuda_fail(e, "mmc_gibbs_create", __FILE__, __LINE__);
        mmc_gibbs_destroy(h);
        return code;
    }
    *out = h;
    return MMC_OK;
}

int mmc_gibbs_set_seed(mmc_gibbs *h, uint64_t seed) {
    MMC_REQUIRE(h, "null handle");
    h->seed = seed;
    h->step = 0;
    return MMC_OK;
}

int mmc_gibbs_set_chain_offset(mmc_gibbs *h, int64_t offset) {
    MMC_REQUIRE(h && offset >= 0, "mmc_gibbs_set_chain_offset: bad arguments");
    h->chain_offset = offset;
    return MMC_OK;
}

int mmc_gibbs_set_out_pitch(mmc_gibbs *h, int64_t pitch_steps) {
    MMC_REQUIRE(h && pitch_steps >= 0, "mmc_gibbs_set_out_pitch: bad arguments");
    h->out_pitch = pitch_steps;
    return MMC_OK;
}

int mmc_gibbs_run_dev(mmc_gibbs *h, int64_t n_collect, int64_t n_discard, double *out_dev, const mmc_replay_gibbs *rp, void *stream) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0 && (out_dev || n_collect == 0), "mmc_gibbs_run_dev: bad arguments");
    MMC_REQUIRE(h->out_pitch == 0 || h->out_pitch >= n_collect, "mmc_gibbs_run_dev: out pitch %lld < n_collect", (long long)h->out_pitch);
    const bool replay = rp && rp->normals && rp->unifs;
    MMC_REQUIRE(!rp || replay || rp->trace, "Gibbs replay needs both the normals and the uniforms tape");
    GibbsParams p{};
    p.state = h->d_state;
    p.out = out_dev;
    p.normals = replay ? rp->normals : nullptr;
    p.unifs = replay ? rp->unifs : nullptr;
    p.trace = rp ? rp->trace : nullptr;
    p.chains = h->chains;
    p.chain_offset = h->chain_offset;
    p.step_base = h->step;
    p.n_collect = n_collect;
    p.n_discard = n_discard;
    p.out_pitch = h->out_pitch > 0 ? h->out_pitch : n_collect;
    p.kind = h->cond.kind;
    p.D = h->dim;
    for (int i = 0; i < 8; ++i) p.p[i] = h->cond.params[i];
    p.key = seed_key(h->seed);
    const unsigned grid = (unsigned)((h->chains + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    if (h->cond.kind >= MMC_G_CUSTOM_BASE) {
        MMC_REQUIRE(!replay && !p.trace, "replay tapes are only defined for the built-in mixture conditional");
        mmc_gibbs_launch_fn fn;
        {
            std::lock_guard<std::mutex> lock(g_cond_mutex);
            fn = cond_registry()[(size_t)(h->cond.kind - MMC_G_CUSTOM_BASE)].fn;
        }
        int rc = fn(&p, h->cond.params, stream);
        if (rc) return rc;
        h->step += n_collect + n_discard;
        return MMC_OK;
    }
    if (h->cond.kind == MMC_G_CONSTANT) gibbs_constant_kernel<<<grid, 128, 0, s>>>(p);
    else if (replay) gibbs_mixture_kernel<true><<<grid, 128, 0, s>>>(p);
    else gibbs_mixture_kernel<false><<<grid, 128, 0, s>>>(p);
    MMC_CUDA(cudaGetLastError());
    h->step += n_collect + n_discard;
    return MMC_OK;
}

int mmc_gibbs_run(mmc_gibbs *h, int64_t n_collect, int64_t n_discard, double *out_host, const mmc_replay_gibbs *replay) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0 && (out_host || n_collect == 0), "mmc_gibbs_run: bad arguments");
    MMC_REQUIRE(h->out_pitch == 0, "mmc_gibbs_run: an output pitch only applies to mmc_gibbs_run_dev");
    const int64_t steps = n_collect + n_discard;
    const size_t out_bytes = (size_t)h->chains * n_collect * h->dim * sizeof(double);
    int rc = grow(&h->d_out, &h->d_out_bytes, out_bytes ? out_bytes : 8);
    if (rc) return rc;
    mmc_replay_gibbs dev{};
    const mmc_replay_gibbs *rp = nullptr;
    if (replay) {
        const size_t tb = (size_t)h->chains * steps * sizeof(double);
        if (replay->normals && replay->unifs) {
            if ((rc = grow(&h->d_tape[0], &h->d_tape_bytes[0], tb ? tb : 8))) return rc;
            if ((rc = grow(&h->d_tape[1], &h->d_tape_bytes[1], tb ? tb : 8))) return rc;
            MMC_CUDA(cudaMemcpyAsync(h->d_tape[0], replay->normals, tb, cudaMemcpyHostToDevice, h->stream));
            MMC_CUDA(cudaMemcpyAsync(h->d_tape[1], replay->unifs, tb, cudaMemcpyHostToDevice, h->stream));
            dev.normals = h->d_tape[0];
            dev.unifs = h->d_tape[1];
        }
        if (replay->trace) {
            if ((rc = grow(&h->d_tape[2], &h->d_tape_bytes[2], 2 * tb ? 2 * tb : 8))) return rc;
            dev.trace = h->d_tape[2];
        }
        rp = &dev;
    }
    rc = mmc_gibbs_run_dev(h, n_collect, n_discard, h->d_out, rp, h->stream);
    if (rc) return rc;
    if (out_bytes) MMC_CUDA(cudaMemcpyAsync(out_host, h->d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (replay && replay->trace)
        MMC_CUDA(cudaMemcpyAsync(replay->trace, h->d_tape[2], (size_t)h->chains * steps * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MMC_CUDA(cudaStreamSynchronize(h->stream));
    return MMC_OK;
}

int mmc_gibbs_run_progress(mmc_gibbs *h, int64_t n_collect, int64_t n_discard, double *out_host, int64_t block, mmc_progress_fn cb,
                           void *user, mmc_run_stats *stats) {
    MMC_REQUIRE(h && n_collect >= 0 && n_discard >= 0 && (out_host || n_collect == 0), "mmc_gibbs_run_progress: bad arguments");
    MMC_REQUIRE(h->out_pitch == 0, "mmc_gibbs_run_progress: an output pitch is set on this handle");
    ProgressSpec sp{h->chains, h->dim, MMC_F64, MMC_TRACK_PER_CHAIN, true, h->d_state};
    auto run_block = [&](int64_t k, void *dst, int64_t pitch, bool) {
        h->out_pitch = pitch;
        const int rc = mmc_gibbs_run_dev(h, k, 0, static_cast<double *>(dst), nullptr, h->stream);
        h->out_pitch = 0;
        return rc;
    };
    auto discard = [&](int64_t k) { return mmc_gibbs_run_dev(h, 0, k, nullptr, nullptr, h->stream); };
    return run_progress_blocks(sp, n_collect, n_discard, out_host, block, cb, user, stats, h->stream, run_block, discard);
}

int mmc_register_gibbs_conditional(const char *name, int32_t dim, mmc_gibbs_launch_fn fn) {
    MMC_REQUIRE(name && fn && dim > 0, "mmc_register_gibbs_conditional: bad arguments");
    std::lock_guard<std::mutex> lock(g_cond_mutex);
    auto &r = cond_registry();
    for (size_t i = 0; i < r.size(); ++i)
        if (r[i].name == name) {
            r[i].dim = dim;
            r[i].fn = fn;
            return MMC_G_CUSTOM_BASE + (int)i;
        }
    r.push_back({name, dim, fn});
    return MMC_G_CUSTOM_BASE + (int)r.size() - 1;
}

int mmc_gibbs_get_state(mmc_gibbs *h, double *state_host) {
    MMC_REQUIRE(h && state_host, "mmc_gibbs_get_state: bad arguments");
    MMC_CUDA(cudaMemcpy(state_host, h->d_state, (size_t)h->chains * h->dim * sizeof(double), cudaMemcpyDeviceToHost));
    return MMC_OK;
}

}  // extern "C"
