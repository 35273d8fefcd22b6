"""Registration hook for custom device targets (include/minimcmc_target.cuh): a user functor compiled with nvcc
into its own shared library runs through the same fused HMC kernel as the built-ins."""
import os
import shutil
import subprocess
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = textwrap.dedent(r"""
    #include "minimcmc_target.cuh"
    // anisotropic Gaussian in 3-D: logp = -0.5 * sum_i x_i^2 / s_i^2, params = s_0, s_1, s_2
    template <class A>
    struct AnisoGauss3 {
        static constexpr int kDim = 3;
        float inv_var[3];
        __host__ explicit AnisoGauss3(const double *p) { for (int i = 0; i < 3; ++i) inv_var[i] = (float)(1.0 / (p[i] * p[i])); }
        __device__ float logp_grad(const float (&x)[3], float (&g)[3]) const {
            float acc = 0.f;
            for (int i = 0; i < 3; ++i) {
                const float t = A::mul(x[i], inv_var[i]);
                acc = A::mad(A::mul(t, x[i]), 0.5f, acc);
                g[i] = -t;
            }
            return -acc;
        }
    };
    MMC_REGISTER_HMC_TARGET(aniso_gauss3, AnisoGauss3)
""")


def test_custom_hmc_target(cuda_device, tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    import mini_mcmc_b200 as mm

    cu = tmp_path / "aniso.cu"
    cu.write_text(SRC)
    so = tmp_path / "libaniso.so"
    lib_dir = os.path.join(ROOT, "mini_mcmc_b200")
    subprocess.run([nvcc, "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), str(cu),
                    "-L", lib_dir, "-l:libminimcmc.so", f"-Xlinker=-rpath,{lib_dir}", "-o", str(so)], check=True)
    sig = (0.5, 1.0, 2.0)
    tgt = mm.CustomTarget(str(so), "aniso_gauss3", 3, sig)
    assert tgt.kind >= 1000
    chains = 2048
    rng = np.random.default_rng(0)
    init = rng.normal(size=(chains, 3)).astype(np.float32)
    # single transition under replay against a numpy restatement
    mom = rng.normal(size=(1, chains, 3)).astype(np.float32)
    u = rng.random((1, chains)).astype(np.float32)
    eps, L = 0.1, 8
    inv_var = (1.0 / np.square(np.array(sig))).astype(np.float32)
    x, p = init.astype(np.float64), mom[0].astype(np.float64)
    grad = lambda z: -z * inv_var
    logp = lambda z: -0.5 * (z * z * inv_var).sum(axis=1)
    h0 = -logp(x) + 0.5 * (p * p).sum(axis=1)
    gh = grad(x) * (eps * 0.5)
    for _ in range(L):
        p = p + gh
        x = x + p * eps
        gh = grad(x) * (eps * 0.5)
        p = p + gh
    h1 = -logp(x) + 0.5 * (p * p).sum(axis=1)
    acc = (h0 - h1) >= np.log(np.maximum(u[0], 1e-38))
    exp = np.where(acc[:, None], x, init)
    hmc = mm.HMC(tgt, init, eps, L)
    got = hmc.run(1, 0, replay=dict(momenta=mom, u=u))[:, 0]
    margin = np.abs((h0 - h1) - np.log(np.maximum(u[0], 1e-38)))
    ok = np.isclose(got, exp, rtol=1e-4, atol=1e-4).all(axis=1) | (margin < 1e-3)
    assert ok.all()
    # long native run: marginal standard deviations
    # (eps * L is kept away from a half period pi * sigma of every coordinate: no resonance)
    s = mm.HMC(tgt, init, 0.11, 7).set_seed(3).run(300, 100).reshape(-1, 3)
    np.testing.assert_allclose(s.std(axis=0), sig, rtol=0.05)


GIBBS_SRC = textwrap.dedent(r"""
    #include "minimcmc_target.cuh"
    // bivariate standard normal with correlation rho: x_i | x_j ~ N(rho x_j, 1 - rho^2); params = rho
    struct BiNormal {
        static constexpr int kDim = 2;
        double rho, sd;
        __host__ explicit BiNormal(const double *p) : rho(p[0]), sd(sqrt(1.0 - p[0] * p[0])) {}
        __device__ double sample(int i, const double (&s)[2], mmc::GibbsRng &rng) const {
            return rho * s[1 - i] + sd * rng.normal();
        }
    };
    MMC_REGISTER_GIBBS_CONDITIONAL(binormal, BiNormal)
""")


def _compile(tmp_path, name, src):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    cu = tmp_path / f"{name}.cu"
    cu.write_text(src)
    so = tmp_path / f"lib{name}.so"
    lib_dir = os.path.join(ROOT, "mini_mcmc_b200")
    subprocess.run([nvcc, "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), str(cu),
                    "-L", lib_dir, "-l:libminimcmc.so", f"-Xlinker=-rpath,{lib_dir}", "-o", str(so)], check=True)
    return str(so)


def test_custom_gibbs_conditional(cuda_device, tmp_path):
    """`Conditional<S>` is user code in the reference (src/distributions.rs:485-487): a device functor registered through
    MMC_REGISTER_GIBBS_CONDITIONAL runs through GibbsSampler like the built-ins."""
    import mini_mcmc_b200 as mm

    so = _compile(tmp_path, "binormal", GIBBS_SRC)
    rho = 0.8
    cond = mm.CustomConditional(so, "binormal", (rho,))
    assert cond.kind >= 1000
    chains = 512
    init = np.zeros((chains, 2))
    s = mm.GibbsSampler(cond, init).set_seed(5)
    x = s.run(400, 100)
    flat = x.reshape(-1, 2)
    np.testing.assert_allclose(flat.mean(axis=0), [0.0, 0.0], atol=0.02)
    np.testing.assert_allclose(flat.std(axis=0), [1.0, 1.0], rtol=0.02)
    np.testing.assert_allclose(np.corrcoef(flat.T)[0, 1], rho, atol=0.01)
    # Philox keyed by (seed, global chain, step, coordinate): shards and continuation reproduce the run
    part = mm.GibbsSampler(cond, init[256:]).set_seed(5).set_chain_offset(256)
    got = np.concatenate([part.run(0, 100), part.run(150, 0), part.run(250, 0)], axis=1)
    np.testing.assert_array_equal(got, x[256:])
    sample, stats = mm.GibbsSampler(cond, init).set_seed(5).run_progress(400, 100, progress=False, block=64)
    np.testing.assert_array_equal(sample, x)
    with pytest.raises(Exception):
        mm.GibbsSampler(cond, np.zeros((4, 3)))     # wrong dimension for the registered conditional


MH_SRC = textwrap.dedent(r"""
    #include "minimcmc_target.cuh"
    // banana-shaped 2-D target: logp = -0.5 * (x0^2 / s^2 + (x1 - b x0^2)^2), params = s, b
    struct Banana {
        static constexpr int kDim = 2;
        double s, b;
        __host__ explicit Banana(const double *p) : s(p[0]), b(p[1]) {}
        __device__ double unnorm_logp(const double (&x)[2]) const {
            const double t = x[1] - b * x[0] * x[0];
            return -0.5 * (x[0] * x[0] / (s * s) + t * t);
        }
    };
    MMC_REGISTER_MH_TARGET(banana, Banana)

    // asymmetric proposal: multiplicative drift towards the origin, y = a x + std z (a = 0.9); param = std
    struct DriftProposal {
        double std;
        __host__ explicit DriftProposal(double p) : std(p) {}
        template <int D> __device__ void sample(const double (&cur)[D], const double (&z)[D], double (&out)[D]) const {
            for (int i = 0; i < D; ++i) out[i] = 0.9 * cur[i] + std * z[i];
        }
        template <int D> __device__ double logp(const double (&from)[D], const double (&to)[D]) const {
            double lp = 0.0;
            for (int i = 0; i < D; ++i) { const double d = to[i] - 0.9 * from[i]; lp += -(d * d) / (2.0 * std * std); }
            return lp;
        }
    };
    MMC_REGISTER_MH_PAIR(banana_drift, Banana, DriftProposal)
""")


def test_custom_mh_target_and_proposal(cuda_device, tmp_path):
    """Any Target / Proposal in MetropolisHastings::new (src/metropolis_hastings.rs:149-159, traits
    src/distributions.rs:92-108): user functors run MHMarkovChain::step (:303-315) on the device."""
    import mini_mcmc_b200 as mm

    so = _compile(tmp_path, "banana", MH_SRC)
    s_, b_ = 1.5, 0.5
    rng = np.random.default_rng(0)
    chains, steps = 400, 6
    init = rng.normal(size=(chains, 2))
    noise = rng.normal(size=(chains, steps, 2))
    u = rng.random((chains, steps))

    def logp(x):
        t = x[:, 1] - b_ * x[:, 0] * x[:, 0]
        return -0.5 * (x[:, 0] * x[:, 0] / (s_ * s_) + t * t)

    def replay(sample, qlogp):
        x = init.copy()
        out = np.empty((chains, steps, 2))
        for s in range(steps):
            y = sample(x, noise[:, s])
            r = (logp(y) + qlogp(y, x)) - (logp(x) + qlogp(x, y))
            acc = r > np.log(u[:, s])
            x = np.where(acc[:, None], y, x)
            out[:, s] = x
        return out

    # built-in IsotropicGaussian proposal with the custom target
    tgt = mm.CustomTarget(so, "banana", 2, (s_, b_), samplers=("mh",))
    std = 0.7
    got = mm.MetropolisHastings(tgt, mm.IsotropicGaussian(std), init).run(steps, 0, replay=dict(noise=noise, u=u))
    exp = replay(lambda x, z: (0.0 + std * z) + x, lambda a, b: (-(b - a) ** 2 / (2 * std * std)).sum(axis=1))
    np.testing.assert_allclose(got, exp, rtol=1e-10, atol=1e-12)
    # custom pair
    pair = mm.CustomTarget(so, "banana_drift", 2, (s_, b_), samplers=("mh",))
    got = mm.MetropolisHastings(pair, mm.CustomProposal(std), init).run(steps, 0, replay=dict(noise=noise, u=u))
    exp = replay(lambda x, z: 0.9 * x + std * z, lambda a, b: (-(b - 0.9 * a) ** 2 / (2 * std * std)).sum(axis=1))
    np.testing.assert_allclose(got, exp, rtol=1e-10, atol=1e-12)
    # native run: x0 ~ N(0, s^2) marginally, E[x1] = b s^2
    smp = mm.MetropolisHastings(pair, mm.CustomProposal(1.0), init).seed(7).run(1500, 500).reshape(-1, 2)
    assert abs(smp[:, 0].std() - s_) < 0.1 and abs(smp[:, 1].mean() - b_ * s_ * s_) < 0.15
    with pytest.raises(Exception):
        mm.MetropolisHastings(pair, mm.CustomProposal(1.0), np.zeros((4, 3)))


NUTS_SRC = textwrap.dedent(r"""
    #include "minimcmc_target.cuh"
    // anisotropic Gaussian in 5-D: logp = -0.5 * sum_i x_i^2 / s_i^2, params = s_0 .. s_4
    template <class A>
    struct AnisoGauss5 {
        static constexpr int kDim = 5;
        float inv_var[5];
        __host__ explicit AnisoGauss5(const double *p) { for (int i = 0; i < 5; ++i) inv_var[i] = (float)(1.0 / (p[i] * p[i])); }
        __device__ float logp_grad(const float (&x)[5], float (&g)[5]) const {
            float acc = 0.f;
            for (int i = 0; i < 5; ++i) {
                const float t = A::mul(x[i], inv_var[i]);
                acc = A::mad(A::mul(t, x[i]), 0.5f, acc);
                g[i] = -t;
            }
            return -acc;
        }
    };
    MMC_REGISTER_HMC_TARGET(aniso_gauss5, AnisoGauss5)
    MMC_REGISTER_NUTS_TARGET(aniso_gauss5, AnisoGauss5)
""")


def test_custom_nuts_target(cuda_device, tmp_path):
    """Any GradientTarget in NUTS::new (src/nuts.rs:123-129, trait src/distributions.rs:78-88): the thread-form functor
    registered for HMC also runs through the one-chain-per-warp NUTS tree kernel, under one kind id."""
    import mini_mcmc_b200 as mm

    so = _compile(tmp_path, "aniso5", NUTS_SRC)
    sig = (0.5, 1.0, 2.0, 1.5, 0.8)
    tgt = mm.CustomTarget(so, "aniso_gauss5", 5, sig, samplers=("hmc", "nuts"))
    chains = 256
    init = np.random.default_rng(1).normal(size=(chains, 5)).astype(np.float32)
    nuts = mm.NUTS(tgt, init, 0.8).set_seed(11)
    s = nuts.run(400, 300)
    assert s.shape == (chains, 400, 5) and np.isfinite(s).all()
    flat = s.reshape(-1, 5)
    np.testing.assert_allclose(flat.std(axis=0), sig, rtol=0.06)
    np.testing.assert_allclose(flat.mean(axis=0), 0.0, atol=0.08)
    # shards + chain offsets reproduce the run (Philox keyed by the global chain)
    part = mm.NUTS(tgt, init[100:], 0.8).set_seed(11).set_chain_offset(100).run(400, 300)
    np.testing.assert_array_equal(part, s[100:])
    # the same kind id serves HMC
    h = mm.HMC(tgt, init, 0.1, 10).set_seed(2).run(200, 100).reshape(-1, 5)
    np.testing.assert_allclose(h.std(axis=0), sig, rtol=0.08)
    with pytest.raises(Exception):
        mm.NUTS(tgt, np.zeros((4, 3), dtype=np.float32), 0.8)
