// minimcmc.hpp — header-only C++ host side over the C ABI (include/minimcmc.h), mirroring the reference crate's
// front-ends with the same names, argument meaning and error behaviour: MetropolisHastings::new/.seed/run/run_progress,
// HMC::new/set_seed/step/run/run_progress, NUTS::new/set_seed/run/run_progress, GibbsSampler::new/set_seed/run/
// run_progress, Categorical, split_rhat_mean_ess, RunStats, init/init_det/init_with_seed, io::save_csv
// (src/metropolis_hastings.rs, src/hmc.rs, src/nuts.rs, src/gibbs.rs, src/stats.rs, src/core.rs, src/io/csv.rs).
// Errors of the library surface as mmc::Error (the crate's Result::Err / panics).
#pragma once

#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "minimcmc.h"

namespace mmc {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc < 0) throw Error(rc, mmc_last_error());
}

// samples are row-major [chains, n_collect, dim]
template <class T>
struct Sample {
    int64_t chains = 0, n_collect = 0, dim = 0;
    std::vector<T> data;
    T &at(int64_t c, int64_t i, int64_t d) { return data[(c * n_collect + i) * dim + d]; }
    const T &at(int64_t c, int64_t i, int64_t d) const { return data[(c * n_collect + i) * dim + d]; }
};

inline std::vector<double> init_with_seed(int64_t n, int64_t d, uint64_t seed) {
    std::vector<double> out((size_t)(n * d));
    check(mmc_init_positions(out.data(), n, d, seed));
    return out;
}
inline std::vector<double> init_det(int64_t n, int64_t d) { return init_with_seed(n, d, 42); }

// RunStats { ess, rhat } with the crate's Display format (src/stats.rs:339-392)
struct RunStats {
    mmc_run_stats raw{};
    static std::string line(const char *name, const mmc_basic_stats &b) {
        char buf[160];
        std::snprintf(buf, sizeof(buf), "%s in [%.2f, %.2f], median: %.2f, mean: %.2f \xc2\xb1 %.2f", name, b.min, b.max, b.median, b.mean, b.std);
        return buf;
    }
    std::string to_string() const { return line("ESS", raw.ess) + "\n" + line("Split R-hat", raw.rhat); }
};

// progress callback of run_progress: (steps done, total, p(accept), max(rhat)) once per block of steps
using ProgressFn = std::function<void(int64_t, int64_t, float, float)>;
namespace detail {
inline void progress_trampoline(int64_t done, int64_t total, float p_accept, float max_rhat, void *user) {
    (*static_cast<ProgressFn *>(user))(done, total, p_accept, max_rhat);
}
}  // namespace detail

// split_rhat_mean_ess(sample [chains, n, dim] f32) -> (rhat[dim], ess[dim]), src/stats.rs:416-423
inline std::pair<std::vector<float>, std::vector<float>> split_rhat_mean_ess(const Sample<float> &s) {
    std::vector<float> rhat((size_t)s.dim), ess((size_t)s.dim);
    check(mmc_split_rhat_ess(s.data.data(), s.chains, s.n_collect, s.dim, rhat.data(), ess.data()));
    return {rhat, ess};
}

// The same over chains that live on several GPUs (one process per GPU): rank 0 obtains the 128-byte id, every rank
// creates the communicator on its own device and calls split_rhat_mean_ess with its local chains in DEVICE memory.
class Communicator {
  public:
    static std::vector<unsigned char> unique_id() {
        std::vector<unsigned char> id(128);
        check(mmc_comm_unique_id(id.data()));
        return id;
    }
    Communicator(const std::vector<unsigned char> &id, int nranks, int rank) { check(mmc_comm_create(&c_, id.data(), nranks, rank)); }
    ~Communicator() { mmc_comm_destroy(c_); }
    Communicator(const Communicator &) = delete;
    std::pair<std::vector<float>, std::vector<float>> split_rhat_mean_ess(const float *sample_dev, int64_t c_local, int64_t n, int64_t p,
                                                                          void *stream = nullptr) {
        std::vector<float> rhat((size_t)p), ess((size_t)p);
        check(mmc_split_rhat_ess_sharded(sample_dev, c_local, n, p, c_, stream, rhat.data(), ess.data()));
        return {rhat, ess};
    }
  private:
    mmc_comm *c_ = nullptr;
};

// io::csv::save_csv, src/io/csv.rs:47-77
template <class T>
inline void save_csv(const Sample<T> &s, const std::string &filename) {
    const int dt = std::is_same<T, float>::value ? MMC_F32 : (std::is_integral<T>::value ? MMC_U64 : MMC_F64);
    static_assert(std::is_same<T, float>::value || sizeof(T) == 8, "samples are f32, f64 or u64");
    check(mmc_save_csv(s.data.empty() ? nullptr : s.data.data(), dt, s.chains, s.n_collect, (int32_t)s.dim, filename.c_str()));
}

inline mmc_proposal_desc isotropic_gaussian(double std_dev) { mmc_proposal_desc q{}; q.kind = MMC_Q_ISO_GAUSSIAN; q.param = std_dev; return q; }
inline mmc_proposal_desc nonnegative_proposal() { mmc_proposal_desc q{}; q.kind = MMC_Q_NONNEG_RW; return q; }
inline mmc_conditional_desc constant_conditional(double c) { mmc_conditional_desc d{}; d.kind = MMC_G_CONSTANT; d.params[0] = c; return d; }
inline mmc_conditional_desc mixture_conditional(double mu0, double sigma0, double mu1, double sigma1, double pi0) {
    mmc_conditional_desc d{};
    d.kind = MMC_G_MIXTURE2;
    d.params[0] = mu0; d.params[1] = sigma0; d.params[2] = mu1; d.params[3] = sigma1; d.params[4] = pi0;
    return d;
}

inline mmc_target_desc target(int kind, int dim, std::initializer_list<double> params = {}) {
    mmc_target_desc t{};
    t.kind = kind;
    t.dim = dim;
    int i = 0;
    for (double p : params) t.params[i++] = p;
    return t;
}

template <class S>
class MetropolisHastings {
  public:
    MetropolisHastings(const mmc_target_desc &t, const mmc_proposal_desc &q, const std::vector<S> &init, int64_t chains, int dim)
        : chains_(chains), dim_(dim) {
        // state type S: f64, f32 (MetropolisHastings<f32, f32, ..>, src/metropolis_hastings.rs:87) or u64 (`usize`)
        static_assert(std::is_same<S, double>::value || std::is_same<S, float>::value || (std::is_integral<S>::value && sizeof(S) == 8),
                      "MetropolisHastings state is f64, f32 or u64");
        check(mmc_mh_create(&h_, &t, &q, init.data(), chains, dim,
                            std::is_integral<S>::value ? MMC_U64 : (std::is_same<S, float>::value ? MMC_F32 : MMC_F64)));
    }
    // Any integer-state target tabulated on [0, logp.size()) with the nonnegative or the reflecting +-1 walk: the `i32`
    // PoissonDist / BinomialDist variants of tests/metrohast_poisson_test.rs:18-85,157-214 (mmc_mh_create_tabulated)
    MetropolisHastings(const std::vector<double> &logp, int proposal_kind, const std::vector<S> &init)
        : chains_((int64_t)init.size()), dim_(1) {
        static_assert(std::is_integral<S>::value && sizeof(S) == 8, "tabulated targets have a u64 state");
        check(mmc_mh_create_tabulated(&h_, logp.data(), (int32_t)logp.size(), proposal_kind, init.data(), chains_));
    }
    // MetropolisHastings::new(Categorical::new(probs), NonnegativeProposal, init), src/distributions.rs:422-477
    MetropolisHastings(const std::vector<double> &probs, const std::vector<S> &init) : chains_((int64_t)init.size()), dim_(1) {
        static_assert(std::is_integral<S>::value && sizeof(S) == 8, "the Categorical target has a u64 state");
        check(mmc_mh_create_categorical(&h_, probs.data(), (int32_t)probs.size(), init.data(), chains_));
    }
    ~MetropolisHastings() { mmc_mh_destroy(h_); }
    MetropolisHastings(const MetropolisHastings &) = delete;
    MetropolisHastings &seed(uint64_t s) { check(mmc_mh_seed(h_, s)); return *this; }
    Sample<S> run(int64_t n_collect, int64_t n_discard) {
        Sample<S> s{chains_, n_collect, dim_, std::vector<S>((size_t)(chains_ * n_collect * dim_))};
        check(mmc_mh_run(h_, n_collect, n_discard, s.data.data(), nullptr));
        return s;
    }
    // ChainRunner::run_progress, src/core.rs:208-360
    std::pair<Sample<S>, RunStats> run_progress(int64_t n_collect, int64_t n_discard, ProgressFn progress = {}, int64_t block = 0) {
        Sample<S> s{chains_, n_collect, dim_, std::vector<S>((size_t)(chains_ * n_collect * dim_))};
        RunStats st;
        check(mmc_mh_run_progress(h_, n_collect, n_discard, s.data.data(), block, progress ? detail::progress_trampoline : nullptr,
                                  progress ? &progress : nullptr, &st.raw));
        return {std::move(s), st};
    }
  private:
    mmc_mh *h_ = nullptr;
    int64_t chains_;
    int dim_;
};

class HMC {
  public:
    HMC(const mmc_target_desc &t, const std::vector<float> &init, int64_t chains, int dim, double step_size, int n_leapfrog)
        : chains_(chains), dim_(dim) { check(mmc_hmc_create(&h_, &t, init.data(), chains, dim, step_size, n_leapfrog)); }
    ~HMC() { mmc_hmc_destroy(h_); }
    HMC(const HMC &) = delete;
    HMC &set_seed(uint64_t s) { check(mmc_hmc_set_seed(h_, s)); return *this; }
    void step() { check(mmc_hmc_step(h_)); }
    Sample<float> run(int64_t n_collect, int64_t n_discard) {
        Sample<float> s{chains_, n_collect, dim_, std::vector<float>((size_t)(chains_ * n_collect * dim_))};
        check(mmc_hmc_run(h_, n_collect, n_discard, s.data.data(), nullptr));
        return s;
    }
    // HMC::run_progress, src/hmc.rs:222-294
    std::pair<Sample<float>, RunStats> run_progress(int64_t n_collect, int64_t n_discard, ProgressFn progress = {}, int64_t block = 0) {
        Sample<float> s{chains_, n_collect, dim_, std::vector<float>((size_t)(chains_ * n_collect * dim_))};
        RunStats st;
        check(mmc_hmc_run_progress(h_, n_collect, n_discard, s.data.data(), block, progress ? detail::progress_trampoline : nullptr,
                                   progress ? &progress : nullptr, &st.raw));
        return {std::move(s), st};
    }
    std::vector<float> positions() {
        std::vector<float> p((size_t)(chains_ * dim_));
        check(mmc_hmc_get_positions(h_, p.data()));
        return p;
    }
  private:
    mmc_hmc *h_ = nullptr;
    int64_t chains_;
    int dim_;
};

class NUTS {
  public:
    NUTS(const mmc_target_desc &t, const std::vector<float> &init, int64_t chains, int dim, double target_accept_p,
         mmc_dtype scalar = MMC_F32, int max_depth = 10)
        : chains_(chains), dim_(dim) { check(mmc_nuts_create(&h_, &t, init.data(), chains, dim, target_accept_p, scalar, max_depth)); }
    ~NUTS() { mmc_nuts_destroy(h_); }
    NUTS(const NUTS &) = delete;
    NUTS &set_seed(uint64_t s) { check(mmc_nuts_set_seed(h_, s)); return *this; }
    // kernel layout (0 = automatic, 32 = one chain per warp, else the lanes per chain compiled in for the target) and
    // work-item slicing of the several-chains-per-warp kernel; neither changes what is sampled
    NUTS &set_layout(int lanes_per_chain) { check(mmc_nuts_set_layout(h_, lanes_per_chain)); return *this; }
    NUTS &set_slicing(int64_t slice_steps) { check(mmc_nuts_set_slicing(h_, slice_steps)); return *this; }
    NUTS &set_regroup(int mode) { check(mmc_nuts_set_regroup(h_, mode)); return *this; }
    int lanes_per_chain() { int32_t n = 0; check(mmc_nuts_get_layout(h_, &n)); return n; }
    Sample<float> run(int64_t n_collect, int64_t n_discard) { return run_impl(n_collect, n_discard, 0); }
    // NUTS::run_progress, src/nuts.rs:194-338 (n_collect + n_discard steps)
    std::pair<Sample<float>, RunStats> run_progress(int64_t n_collect, int64_t n_discard, ProgressFn progress = {}, int64_t block = 0) {
        Sample<float> s{chains_, n_collect, dim_, std::vector<float>((size_t)(chains_ * n_collect * dim_))};
        RunStats st;
        check(mmc_nuts_run_progress(h_, n_collect, n_discard, s.data.data(), block, progress ? detail::progress_trampoline : nullptr,
                                    progress ? &progress : nullptr, &st.raw));
        return {std::move(s), st};
    }
  private:
    Sample<float> run_impl(int64_t n_collect, int64_t n_discard, int progress) {
        Sample<float> s{chains_, n_collect, dim_, std::vector<float>((size_t)(chains_ * n_collect * dim_))};
        check(mmc_nuts_run(h_, n_collect, n_discard, progress, s.data.data(), nullptr));
        return s;
    }
    mmc_nuts *h_ = nullptr;
    int64_t chains_;
    int dim_;
};

// GibbsSampler::new(conditional, initial_states), src/gibbs.rs:165-186
class GibbsSampler {
  public:
    GibbsSampler(const mmc_conditional_desc &cond, const std::vector<double> &init, int64_t chains, int dim) : chains_(chains), dim_(dim) {
        check(mmc_gibbs_create(&h_, &cond, init.data(), chains, dim));
    }
    ~GibbsSampler() { mmc_gibbs_destroy(h_); }
    GibbsSampler(const GibbsSampler &) = delete;
    GibbsSampler &set_seed(uint64_t s) { check(mmc_gibbs_set_seed(h_, s)); return *this; }
    Sample<double> run(int64_t n_collect, int64_t n_discard) {
        Sample<double> s{chains_, n_collect, dim_, std::vector<double>((size_t)(chains_ * n_collect * dim_))};
        check(mmc_gibbs_run(h_, n_collect, n_discard, s.data.data(), nullptr));
        return s;
    }
    std::pair<Sample<double>, RunStats> run_progress(int64_t n_collect, int64_t n_discard, ProgressFn progress = {}, int64_t block = 0) {
        Sample<double> s{chains_, n_collect, dim_, std::vector<double>((size_t)(chains_ * n_collect * dim_))};
        RunStats st;
        check(mmc_gibbs_run_progress(h_, n_collect, n_discard, s.data.data(), block, progress ? detail::progress_trampoline : nullptr,
                                     progress ? &progress : nullptr, &st.raw));
        return {std::move(s), st};
    }
  private:
    mmc_gibbs *h_ = nullptr;
    int64_t chains_;
    int dim_;
};

}  // namespace mmc
