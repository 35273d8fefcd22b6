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
