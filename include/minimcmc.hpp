// minimcmc.hpp — header-only C++ façade over the C ABI (include/minimcmc.h), mirroring the reference crate's
// front-ends: MetropolisHastings::new/.seed/run, HMC::new/set_seed/step/run, NUTS::new/set_seed/run/run_progress,
// split_rhat_mean_ess, init_det (src/metropolis_hastings.rs, src/hmc.rs, src/nuts.rs, src/stats.rs, src/core.rs).
#pragma once

#include <stdexcept>
#include <string>
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
};

inline std::vector<double> init_with_seed(int64_t n, int64_t d, uint64_t seed) {
    std::vector<double> out((size_t)(n * d));
    check(mmc_init_positions(out.data(), n, d, seed));
    return out;
}
inline std::vector<double> init_det(int64_t n, int64_t d) { return init_with_seed(n, d, 42); }

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
        check(mmc_mh_create(&h_, &t, &q, init.data(), chains, dim, sizeof(S) == 8 && std::is_integral<S>::value ? MMC_U64 : MMC_F64));
    }
    ~MetropolisHastings() { mmc_mh_destroy(h_); }
    MetropolisHastings(const MetropolisHastings &) = delete;
    MetropolisHastings &seed(uint64_t s) { check(mmc_mh_seed(h_, s)); return *this; }
    Sample<S> run(int64_t n_collect, int64_t n_discard) {
        Sample<S> s{chains_, n_collect, dim_, std::vector<S>((size_t)(chains_ * n_collect * dim_))};
        check(mmc_mh_run(h_, n_collect, n_discard, s.data.data(), nullptr));
        return s;
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
    Sample<float> run(int64_t n_collect, int64_t n_discard) { return run_impl(n_collect, n_discard, 0); }
    std::pair<Sample<float>, mmc_run_stats> run_progress(int64_t n_collect, int64_t n_discard) {
        Sample<float> s = run_impl(n_collect, n_discard, 1);
        std::vector<float> rhat((size_t)dim_), ess((size_t)dim_);
        check(mmc_split_rhat_ess(s.data.data(), chains_, n_collect, dim_, rhat.data(), ess.data()));
        mmc_run_stats st{};
        check(mmc_basic_stats_of(ess.data(), dim_, &st.ess));
        check(mmc_basic_stats_of(rhat.data(), dim_, &st.rhat));
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

}  // namespace mmc
