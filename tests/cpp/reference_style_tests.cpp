// Host-side tests written against the C++ mirror of the crate API (include/minimcmc.hpp).  Each test restates one of the
// reference's own tests with the same structure and thresholds; the test name cites it.  Built and run by
// tests/test_gpu_cpp_host.py on the GPU box (exit code 0 = all passed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "minimcmc.hpp"

static int g_failed = 0;
#define EXPECT(cond, ...)                                                      \
    do {                                                                       \
        if (!(cond)) {                                                         \
            std::printf("  FAILED %s:%d: %s -- ", __FILE__, __LINE__, #cond);  \
            std::printf(__VA_ARGS__);                                          \
            std::printf("\n");                                                 \
            ++g_failed;                                                        \
        }                                                                      \
    } while (0)

static std::vector<float> to_f32(const std::vector<double> &v) { return std::vector<float>(v.begin(), v.end()); }

// src/gibbs.rs:277-290 test_gibbs_chain_step
static void test_gibbs_chain_step() {
    mmc::GibbsSampler chain(mmc::constant_conditional(7.0), {0.0, 0.0, 0.0}, 1, 3);
    auto s = chain.run(1, 0);
    for (double x : s.data) EXPECT(std::fabs(x - 7.0) < 2.3e-16, "expected 7.0, got %g", x);
}

// src/gibbs.rs:292-305 test_gibbs_sampler_run, :307-325 test_gibbs_sampler_run_progress
static void test_gibbs_sampler_run_and_run_progress() {
    mmc::GibbsSampler sampler(mmc::constant_conditional(42.0), mmc::init_det(4, 2), 4, 2);
    auto sample = sampler.set_seed(42).run(10, 5);
    EXPECT(sample.chains == 4 && sample.n_collect == 10 && sample.dim == 2, "shape");
    for (double x : sample.data) EXPECT(x == 42.0, "expected 42, got %g", x);
    mmc::GibbsSampler sampler2(mmc::constant_conditional(42.0), mmc::init_det(4, 2), 4, 2);
    auto res = sampler2.run_progress(10, 5);
    EXPECT(res.first.data.size() == 80, "shape");
    for (double x : res.first.data) EXPECT(x == 42.0, "expected 42, got %g", x);
    std::printf("  %s\n", res.second.to_string().c_str());
}

// src/gibbs.rs:327-392 assert_mixture_simulation / test_gibbs_sampler_mixture_1
static void test_gibbs_sampler_mixture_1() {
    const double mu0 = -2.0, sigma0 = 1.0, mu1 = 3.0, sigma1 = 1.5, pi0 = 0.5;
    const double theo_mean = pi0 * mu0 + (1.0 - pi0) * mu1;
    const double theo_var = pi0 * (sigma0 * sigma0 + (mu0 - theo_mean) * (mu0 - theo_mean)) +
                            (1.0 - pi0) * (sigma1 * sigma1 + (mu1 - theo_mean) * (mu1 - theo_mean));
    mmc::GibbsSampler sampler(mmc::mixture_conditional(mu0, sigma0, mu1, sigma1, pi0), mmc::init_det(4, 2), 4, 2);
    auto sample = sampler.set_seed(42).run(100000, 10000);
    double sum = 0.0, sq = 0.0;
    const int64_t n = sample.chains * sample.n_collect;
    for (int64_t c = 0; c < 4; ++c)
        for (int64_t i = 0; i < sample.n_collect; ++i) sum += sample.at(c, i, 0);
    const double mean = sum / (double)n;
    for (int64_t c = 0; c < 4; ++c)
        for (int64_t i = 0; i < sample.n_collect; ++i) sq += (sample.at(c, i, 0) - mean) * (sample.at(c, i, 0) - mean);
    const double var = sq / (double)(n - 1);
    EXPECT(std::fabs(mean - theo_mean) < std::fabs(theo_mean) / 10.0, "mean %g vs %g", mean, theo_mean);
    EXPECT(std::fabs(var - theo_var) < std::fabs(theo_var) / 10.0, "var %g vs %g", var, theo_var);
}

// examples/minimal_mh.rs + src/metropolis_hastings.rs:338-380 (2-D Gaussian target, isotropic proposal): sample moments
static void test_mh_gaussian2d_moments() {
    const std::vector<double> init = mmc::init_det(4, 2);
    mmc::MetropolisHastings<double> mh(mmc::target(MMC_T_GAUSSIAN2D, 2, {0.0, 1.0, 4.0, 2.0, 2.0, 3.0}), mmc::isotropic_gaussian(1.0), init, 4, 2);
    auto s = mh.seed(42).run(50000, 5000);
    double m0 = 0, m1 = 0;
    const double n = (double)(s.chains * s.n_collect);
    for (int64_t c = 0; c < 4; ++c)
        for (int64_t i = 0; i < s.n_collect; ++i) { m0 += s.at(c, i, 0); m1 += s.at(c, i, 1); }
    m0 /= n; m1 /= n;
    double c00 = 0, c01 = 0, c11 = 0;
    for (int64_t c = 0; c < 4; ++c)
        for (int64_t i = 0; i < s.n_collect; ++i) {
            const double a = s.at(c, i, 0) - m0, b = s.at(c, i, 1) - m1;
            c00 += a * a; c01 += a * b; c11 += b * b;
        }
    c00 /= n - 1; c01 /= n - 1; c11 /= n - 1;
    EXPECT(std::fabs(m0 - 0.0) < 0.5 && std::fabs(m1 - 1.0) < 0.5, "mean (%g, %g)", m0, m1);
    EXPECT(std::fabs(c00 - 4.0) < 0.5 && std::fabs(c01 - 2.0) < 0.5 && std::fabs(c11 - 3.0) < 0.5, "cov (%g, %g, %g)", c00, c01, c11);
}

// tests/metrohast_poisson_test.rs:90-130 / examples/poisson_mh.rs: Poisson(4) through the +-1 nonnegative walk
static void test_mh_poisson_mean_and_variance() {
    std::vector<uint64_t> init(64, 0);
    mmc::MetropolisHastings<uint64_t> mh(mmc::target(MMC_T_POISSON, 1, {4.0}), mmc::nonnegative_proposal(), init, 64, 1);
    std::vector<std::pair<int64_t, float>> seen;
    auto res = mh.seed(42).run_progress(10000, 1000, [&](int64_t done, int64_t total, float p_accept, float max_rhat) {
        (void)max_rhat;
        EXPECT(done <= total && total == 11000, "progress %lld / %lld", (long long)done, (long long)total);
        seen.push_back({done, p_accept});
    });
    EXPECT(!seen.empty() && seen.back().first == 11000, "progress callback");
    EXPECT(seen.back().second > 0.3f && seen.back().second < 1.0f, "p(accept) %g", seen.back().second);
    double sum = 0, sq = 0;
    const double n = (double)res.first.data.size();
    for (uint64_t k : res.first.data) sum += (double)k;
    const double mean = sum / n;
    for (uint64_t k : res.first.data) sq += ((double)k - mean) * ((double)k - mean);
    EXPECT(std::fabs(mean - 4.0) < 0.2 && std::fabs(sq / (n - 1) - 4.0) < 0.4, "mean %g var %g", mean, sq / (n - 1));
    std::printf("  %s\n", res.second.to_string().c_str());
    // the block-wise run_progress draws equal those of run()
    mmc::MetropolisHastings<uint64_t> mh2(mmc::target(MMC_T_POISSON, 1, {4.0}), mmc::nonnegative_proposal(), init, 64, 1);
    auto plain = mh2.seed(42).run(10000, 1000);
    EXPECT(plain.data == res.first.data, "run_progress draws differ from run()");
}

// src/distributions.rs:422-477 Categorical as an MH target
static void test_mh_categorical_frequencies() {
    std::vector<uint64_t> init(256, 0);
    mmc::MetropolisHastings<uint64_t> mh(std::vector<double>{2.0, 3.0, 5.0}, init);
    auto s = mh.seed(1).run(4000, 400);
    double f[3] = {0, 0, 0};
    for (uint64_t k : s.data) { EXPECT(k < 3, "category %llu", (unsigned long long)k); if (k < 3) f[k] += 1.0; }
    for (int k = 0; k < 3; ++k) f[k] /= (double)s.data.size();
    EXPECT(std::fabs(f[0] - 0.2) < 0.01 && std::fabs(f[1] - 0.3) < 0.01 && std::fabs(f[2] - 0.5) < 0.01, "freq %g %g %g", f[0], f[1], f[2]);
}

// src/hmc.rs:456-573 (test_single / test_3_chains / test_progress: output shapes), examples/rosenbrock3d_hmc.rs
static void test_hmc_shapes_and_run_progress() {
    auto init = to_f32(mmc::init_det(4, 3));
    mmc::HMC hmc(mmc::target(MMC_T_ROSENBROCK_ND, 3), init, 4, 3, 0.01, 10);
    int calls = 0;
    auto res = hmc.set_seed(42).run_progress(400, 50, [&](int64_t done, int64_t total, float, float) { ++calls; EXPECT(done <= total && total == 400, "progress"); });
    EXPECT(res.first.chains == 4 && res.first.n_collect == 400 && res.first.dim == 3, "shape");
    EXPECT(calls > 0, "progress callback never called");
    for (float x : res.first.data) EXPECT(std::isfinite(x), "non-finite draw");
    mmc::HMC hmc2(mmc::target(MMC_T_ROSENBROCK_ND, 3), init, 4, 3, 0.01, 10);
    auto plain = hmc2.set_seed(42).run(400, 50);
    EXPECT(plain.data == res.first.data, "run_progress draws differ from run()");
    hmc2.step();
    EXPECT(hmc2.positions().size() == 12, "positions");
}

// examples/minimal_nuts.rs:9-30 (Rosenbrock2D a = 1, b = 100, 4 chains, 400 + 400) + src/nuts.rs:1206-1221 shapes
static void test_nuts_run_and_run_progress() {
    auto init = to_f32(mmc::init_det(4, 2));
    mmc::NUTS nuts(mmc::target(MMC_T_ROSENBROCK_2D, 2, {1.0, 100.0}), init, 4, 2, 0.95);
    auto res = nuts.set_seed(42).run_progress(400, 400, [](int64_t, int64_t, float, float) {}, 128);
    EXPECT(res.first.chains == 4 && res.first.n_collect == 400 && res.first.dim == 2, "shape");
    for (float x : res.first.data) EXPECT(std::isfinite(x), "non-finite draw");
    EXPECT(res.second.raw.ess.min > 0.0f, "ESS %g", res.second.raw.ess.min);
    mmc::NUTS nuts2(mmc::target(MMC_T_ROSENBROCK_2D, 2, {1.0, 100.0}), init, 4, 2, 0.95);
    auto plain = nuts2.set_seed(42).run_progress(400, 400);      // one block covering... default blocks: still equal draws
    EXPECT(plain.first.data == res.first.data, "run_progress draws depend on the block size");
    auto r = nuts2.run(10, 0);                                   // NUTS::run: slot 0 holds the starting position (src/nuts.rs:460)
    EXPECT(r.n_collect == 10, "shape");
    EXPECT(nuts2.lanes_per_chain() == 4, "2-D targets run 8 chains per warp by default, got %d lanes per chain", nuts2.lanes_per_chain());
    mmc::NUTS warp(mmc::target(MMC_T_ROSENBROCK_2D, 2, {1.0, 100.0}), init, 4, 2, 0.95);
    auto w = warp.set_seed(42).set_layout(32).run_progress(50, 50);   // one chain per warp: same algorithm and counters
    EXPECT(warp.lanes_per_chain() == 32, "layout 32");
    for (float x : w.first.data) EXPECT(std::isfinite(x), "non-finite draw");
}

// src/stats.rs:810-834 ess_1: iid draws give ESS ~ chains * n and split-Rhat ~ 1
static void test_split_rhat_mean_ess_iid() {
    mmc::Sample<float> s{4, 1000, 1, to_f32(mmc::init_with_seed(4000, 1, 42))};
    auto re = mmc::split_rhat_mean_ess(s);
    EXPECT(re.second[0] > 3800.0f && re.second[0] < 4300.0f, "ESS %g", re.second[0]);
    EXPECT(std::fabs(re.first[0] - 1.0f) < 0.01f, "Rhat %g", re.first[0]);
}

// src/io/csv.rs:178-217 test_save_csv_single_chain_single_obs / test_save_csv_multi_chain
static void test_save_csv() {
    auto slurp = [](const std::string &p) { std::ifstream f(p); std::stringstream ss; ss << f.rdbuf(); return ss.str(); };
    const std::string path = "/tmp/mmc_cpp_test.csv";
    mmc::save_csv(mmc::Sample<double>{1, 1, 1, {42.0}}, path);
    EXPECT(slurp(path) == "chain,observation,dim_0\n0,0,42\n", "got %s", slurp(path).c_str());
    mmc::save_csv(mmc::Sample<uint64_t>{2, 2, 2, {1, 2, 3, 4, 10, 20, 30, 40}}, path);
    EXPECT(slurp(path) == "chain,observation,dim_0,dim_1\n0,0,1,2\n0,1,3,4\n1,0,10,20\n1,1,30,40\n", "got %s", slurp(path).c_str());
    std::remove(path.c_str());
}

// error behaviour: invalid arguments surface as mmc::Error, like the crate's Err / panics
static void test_errors() {
    bool threw = false;
    try { mmc::GibbsSampler bad(mmc::mixture_conditional(0, 1, 1, 1, 0.5), {0.0, 0.0, 0.0}, 1, 3); } catch (const mmc::Error &e) { threw = e.code == MMC_ERR_INVALID; }
    EXPECT(threw, "mixture conditional with dim 3 must be rejected");
    threw = false;
    try { mmc::MetropolisHastings<uint64_t> bad(std::vector<double>{0.5, 0.5}, std::vector<uint64_t>{5}); } catch (const mmc::Error &e) { threw = e.code == MMC_ERR_INVALID; }
    EXPECT(threw, "start outside the support must be rejected");
}

// tests/metrohast_poisson_test.rs:157-249: Binomial(10, 0.3) through the +-1 walk clamped to [0, 10], i32 state there
static double ln_factorial(int k) {
    if (k < 2) return 0.0;
    double acc = 0.0;
    for (int i = 1; i <= k; ++i) acc += std::log((double)i);
    return acc;
}
static void test_binomial_mh() {
    const int n = 10;
    const double p = 0.3;
    std::vector<double> logp;
    for (int k = 0; k <= n; ++k)
        logp.push_back((ln_factorial(n) - ln_factorial(k) - ln_factorial(n - k)) + (double)k * std::log(p) + ((double)n - (double)k) * std::log(1.0 - p));
    std::vector<uint64_t> init(32, 5);   // start from the middle
    mmc::MetropolisHastings<uint64_t> mh(logp, MMC_Q_REFLECT_RW, init);
    auto s = mh.seed(42).run(20000, 2000);
    std::vector<double> freq(n + 1, 0.0);
    for (uint64_t k : s.data) {
        EXPECT(k <= (uint64_t)n, "state %llu outside the support", (unsigned long long)k);
        if (k <= (uint64_t)n) freq[k] += 1.0 / (double)s.data.size();
    }
    for (int k = 0; k <= n; ++k) {
        double nck = 1.0;
        for (int i = 1; i <= k; ++i) nck = nck * (double)(n - k + i) / (double)i;
        const double pmf = nck * std::pow(p, k) * std::pow(1.0 - p, n - k);
        EXPECT(std::fabs(freq[k] - pmf) < 0.05, "k = %d: frequency %g, pmf %g", k, freq[k], pmf);
    }
}

// MetropolisHastings<f32, f32, ..> (src/metropolis_hastings.rs:87): the Gaussian2D moments of :338-401 on f32 state
static void test_mh_f32_state() {
    std::vector<float> init(64 * 2, 0.0f);
    mmc::MetropolisHastings<float> mh(mmc::target(MMC_T_GAUSSIAN2D, 2, {0.0, 1.0, 4.0, 2.0, 2.0, 3.0}), mmc::isotropic_gaussian(1.0), init, 64, 2);
    auto s = mh.seed(42).run(2000, 500);
    double m0 = 0, m1 = 0;
    const double cnt = (double)(s.data.size() / 2);
    for (size_t i = 0; i < s.data.size(); i += 2) { m0 += s.data[i]; m1 += s.data[i + 1]; }
    EXPECT(std::fabs(m0 / cnt - 0.0) < 0.3 && std::fabs(m1 / cnt - 1.0) < 0.3, "mean (%g, %g)", m0 / cnt, m1 / cnt);
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "--host-only") {   // CPU boxes: only the pieces that need no device
        test_save_csv();
        std::printf(g_failed ? "FAILED\n" : "ok (host only)\n");
        return g_failed ? 1 : 0;
    }
    struct { const char *name; void (*fn)(); } tests[] = {
        {"test_gibbs_chain_step", test_gibbs_chain_step},
        {"test_gibbs_sampler_run_and_run_progress", test_gibbs_sampler_run_and_run_progress},
        {"test_gibbs_sampler_mixture_1", test_gibbs_sampler_mixture_1},
        {"test_mh_gaussian2d_moments", test_mh_gaussian2d_moments},
        {"test_mh_poisson_mean_and_variance", test_mh_poisson_mean_and_variance},
        {"test_mh_categorical_frequencies", test_mh_categorical_frequencies},
        {"test_binomial_mh", test_binomial_mh},
        {"test_mh_f32_state", test_mh_f32_state},
        {"test_hmc_shapes_and_run_progress", test_hmc_shapes_and_run_progress},
        {"test_nuts_run_and_run_progress", test_nuts_run_and_run_progress},
        {"test_split_rhat_mean_ess_iid", test_split_rhat_mean_ess_iid},
        {"test_save_csv", test_save_csv},
        {"test_errors", test_errors},
    };
    for (auto &t : tests) {
        const int before = g_failed;
        std::printf("%s\n", t.name);
        try { t.fn(); } catch (const mmc::Error &e) { std::printf("  FAILED: mmc::Error %d: %s\n", e.code, e.what()); ++g_failed; }
        std::printf("  %s\n", g_failed == before ? "ok" : "FAILED");
    }
    std::printf("%d failure(s)\n", g_failed);
    return g_failed ? 1 : 0;
}
