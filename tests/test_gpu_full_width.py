"""Parity at BASELINE.json's full widths for C3, C4 and C5 (C2's full-width check lives in tests/test_gpu_mh.py).

A replayed single transition is independent per chain, so the whole-width launch of the PRODUCTION kernels (the grid /
tile / wave shapes the bench numbers come from) can be compared chain by chain with the oracle:
  * C3: 262,144 chains, RosenbrockND D = 3, L = 50 - every chain through the oracle (src/hmc.rs:304-431);
  * C5: 65,536 chains, RosenbrockND D = 100 - every chain through the oracle, one NUTS transition from positions and
    adaptation states the device itself reached after a native warm-up (src/nuts.rs:550-691);
  * C4: 32,768 chains, D = 1,024 dense Gaussian, L = 50 on the default tcgen05 path - the oracle (a CPU matvec per
    leapfrog) checks a spread-out sample of chains that covers every CTA-pair tile row parity and both cluster ranks; the
    rest of the width is covered by a size-independent property: chains are independent, so any window of the launch
    re-run on its own must reproduce its rows bit for bit.
Tolerances are the ones of tests/single_transition.py (RTOL = 1e-5, ties excluded by the oracle's own margins).
"""
import numpy as np
import pytest

import oracle
import single_transition as st
from test_gpu_single_transition import check_hmc

pytestmark = pytest.mark.gpu
RTOL = st.RTOL


@pytest.fixture(scope="module")
def mm(cuda_device):
    import mini_mcmc_b200 as m

    return m


def frac(a, tol):
    return float((np.asarray(a) <= tol).mean())


def test_c3_full_width_single_transition(mm):
    """hmc_run_pair_kernel (and the exact kernel) at the C3 launch shape, 262,144 chains.  Measured distribution of the
    throughput kernel (scripts/parity_probe.py full, profiles/r3_parity_full_width.log): log p(current) <= 7e-7 on every
    chain, accept_logp inside RTOL on 99.997 % (max 2.5e-5), states inside RTOL on 99.8 % (p99.99 4.5e-5, max 1.7e-4),
    no accept decision differs; the exact kernel reproduces the oracle bit for bit.  The tails are the chains whose 50
    leapfrogs run along the stiff Rosenbrock ridge (see test_hmc_c3_single_transition_wide_and_f64_shadow: the f32
    reference is as far from the float64 trajectory as the kernel is from the reference)."""
    case = st.hmc_case(3, chains=262144)
    exp = st.hmc_oracle(case)
    cmp = st.hmc_compare(case, exp, st.hmc_device(mm, case, False))
    assert not cmp["unexplained"].any() and cmp["differ"].mean() <= 1e-3
    assert cmp["logp_cur"].max() <= RTOL
    assert frac(cmp["accept_logp"], RTOL) >= 0.9999 and cmp["accept_logp"].max() <= 1e-4, f"accept_logp {cmp['accept_logp'].max():.2e}"
    assert frac(cmp["state"], RTOL) >= 0.995 and np.quantile(cmp["state"], 0.9999) <= 1e-4 and cmp["state"].max() <= 1e-3, \
        f"state {frac(cmp['state'], RTOL):.5f} {cmp['state'].max():.2e}"
    assert frac(cmp["logp_prop"], RTOL) >= 0.97 and np.quantile(cmp["logp_prop"], 0.999) <= 1e-4, f"logp_prop {frac(cmp['logp_prop'], RTOL):.5f}"
    check_hmc(st.hmc_compare(case, exp, st.hmc_device(mm, case, True)), True, "C3 full width exact")


def test_c5_full_width_single_transition(mm):
    """nuts_group_kernel (4 chains per warp, sliced work items; packed f32x2 for the throughput policy) at the C5 launch
    shape: the device warms 65,536 chains up natively, then device and oracle take the same replayed transition from that
    state.  Measured (profiles/r3_parity_full_width.log): joint_0, log u, alpha <= 5e-7 on every chain; x' inside RTOL on
    99.998 % (max 1.5e-5, all of depth <= 5 below 4e-7); 15 chains take another branch than the oracle, 14 inside the
    oracle's own tie margin (1e-4) and one at a relative distance of 4e-4 from a U-turn threshold in a deep tree."""
    case = st.nuts_full_width_case(mm)
    exp = st.nuts_oracle(case)
    assert exp["trace"][:, 5].max() >= 5 and (exp["trace"][:, 5] >= 4).mean() > 0.5   # the C5 depth mix
    for exact in (True, False):
        got = st.nuts_device(mm, case, 0, exact)
        assert got["lanes"] == 8
        cmp = st.nuts_compare(exp, got)
        what = f"C5 full width exact={exact}"
        un = cmp["unexplained"]
        assert un.sum() <= (0 if exact else 3) and (cmp["margin"][un] < 1e-3).all(), f"{what}: unexplained margins {cmp['margin'][un]}"
        assert cmp["differ"].mean() <= 1e-3
        for k in ("joint", "logu", "eps", "alpha"):
            assert cmp[k].max() <= RTOL, f"{what}: {k} {cmp[k].max():.2e}"
        shallow = cmp["depth"] <= 5
        assert cmp["state"][shallow].max() <= RTOL
        assert frac(cmp["state"], RTOL) >= 0.9999 and cmp["state"].max() <= (1e-6 if exact else 1e-4), f"{what}: state {cmp['state'].max():.2e}"


def test_c4_full_width_sampled_and_windowed(mm):
    chains, D, L, eps = 32768, 1024, 50, 0.05
    rng = np.random.default_rng(42)
    A = rng.normal(size=(D, D)).astype(np.float32)
    cov = (A @ A.T / D + np.eye(D, dtype=np.float32)).astype(np.float64)
    mean = rng.normal(size=D)
    tgt = mm.DenseGaussian(mean, cov)
    init = (rng.standard_normal(size=(chains, D), dtype=np.float32) + mean.astype(np.float32))
    mom = rng.standard_normal(size=(1, chains, D), dtype=np.float32)
    u = rng.random((1, chains), dtype=np.float32)
    h = mm.HMC(tgt, init, eps, L)   # default path: tcgen05 CTA pairs
    trace = np.zeros((1, chains, 4), dtype=np.float32)
    got = h.run(1, 0, replay=dict(momenta=mom, u=u), trace=trace)
    assert np.isfinite(got).all() and 0.5 < trace[0, :, 3].mean() <= 1.0
    # (1) oracle on a sample: first / last tile, both CTAs of a pair (rows 0-127 / 128-255 of a 256-row tile), odd tiles
    idx = np.unique(np.concatenate([np.arange(0, 4), np.arange(126, 130), np.arange(254, 258), rng.integers(0, chains, 48),
                                    np.arange(chains - 4, chains)]))
    otgt = oracle.dense_gaussian(tgt.mean, tgt.precision, tgt.norm_const)
    exp, _, exp_tr = oracle.hmc_run_replay(otgt, init[idx], eps, L, 1, 0, np.ascontiguousarray(mom[:, idx]),
                                           np.ascontiguousarray(u[:, idx]), want_trace=True)
    cmp = st.hmc_compare(dict(u=u[:, idx]), dict(out=exp[:, 0], trace=exp_tr[0]), dict(out=got[idx, 0], trace=trace[0, idx]))
    assert not cmp["unexplained"].any()
    for k in ("logp_cur", "logp_prop", "accept_logp"):
        assert cmp[k].max() <= RTOL, f"{k} {cmp[k].max():.2e}"
    # 50 dense leapfrogs: the tensor core's truncating fp32 accumulation adds ~1e-7 per GEMM, a random walk the FP32 SIMT path
    # does not have (measured 6e-6 median / 1.2e-5 max at L = 50 vs 4e-7; RTOL holds for L <= 7 in tests/test_gpu_dense.py)
    assert np.median(cmp["state"]) <= RTOL and cmp["state"].max() <= 3e-5, f"state {cmp['state'].max():.2e}"
    # (2) independence of the chains: windows of the launch re-run on their own reproduce their rows bit for bit
    for lo, n in ((0, 256), (12800, 512), (chains - 300, 300)):
        part = mm.HMC(tgt, init[lo:lo + n], eps, L)
        tr = np.zeros((1, n, 4), dtype=np.float32)
        sub = part.run(1, 0, replay=dict(momenta=np.ascontiguousarray(mom[:, lo:lo + n]), u=np.ascontiguousarray(u[:, lo:lo + n])),
                       trace=tr)
        # (the quadratic form is summed over the column tiles with float atomics, so log-probs may differ in the last bit
        # between two launches and an exact tie could flip; the trajectories themselves are bit-reproducible)
        same = tr[0, :, 3] == trace[0, lo:lo + n, 3]
        assert same.mean() >= 0.999
        np.testing.assert_array_equal(sub[same], got[lo:lo + n][same])
        np.testing.assert_allclose(tr[0, :, :2], trace[0, lo:lo + n, :2], rtol=2e-6)
