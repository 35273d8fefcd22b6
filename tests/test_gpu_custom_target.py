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
