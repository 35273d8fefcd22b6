"""Gibbs sampler (src/gibbs.rs) on the device against the oracle restatement and the reference's own tests."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mm(cuda_device):
    import mini_mcmc_b200 as m

    return m


def test_constant_conditional_like_reference_tests(mm):
    # test_gibbs_chain_step / test_gibbs_sampler_run / _run_progress, src/gibbs.rs:277-325
    s = mm.GibbsSampler(mm.ConstantConditional(7.0), np.zeros((1, 3)))
    out = s.run(1, 0)
    np.testing.assert_array_equal(out, np.full((1, 1, 3), 7.0))
    s = mm.GibbsSampler(mm.ConstantConditional(42.0), mm.init_det(4, 2)).set_seed(42)
    out = s.run(10, 5)
    assert out.shape == (4, 10, 2)
    np.testing.assert_array_equal(out, np.full((4, 10, 2), 42.0))
    sample, stats = mm.GibbsSampler(mm.ConstantConditional(42.0), mm.init_det(4, 2)).run_progress(10, 5, progress=False)
    np.testing.assert_array_equal(sample, np.full((4, 10, 2), 42.0))


@pytest.mark.parametrize("params", [(-2.0, 1.0, 3.0, 1.5, 0.5), (-42.0, 69.0, 1.0, 2.0, 0.123), (0.0, 0.5, 0.2, 0.6, 0.9)])
def test_mixture_replay_of_reference_stream_matches_oracle(mm, params):
    # the reference's own draws (every chain holds a clone of the conditional's SmallRng(42), src/gibbs.rs:165-176)
    init = mm.init_det(6, 2)
    ref = oracle.gibbs_run(oracle.G_MIXTURE2, params, init, 300, 50, cond_seed=42, record=True)
    normals, unifs = ref["tapes"]
    s = mm.GibbsSampler(mm.MixtureConditional(*params), init)
    out = s.run(300, 50, replay=dict(normals=normals, unifs=unifs))
    # z decisions and x values: f64 with the reference's operation order; exp/sqrt differ from the host libm by <= 1 ulp
    np.testing.assert_array_equal(out[:, :, 1], ref["out"][:, :, 1])
    np.testing.assert_allclose(out[:, :, 0], ref["out"][:, :, 0], rtol=1e-14)
    np.testing.assert_allclose(s.current_state(), ref["state"], rtol=1e-14)


def test_mixture_native_stream_equals_oracle_replay_of_its_tape(mm):
    params = (-2.0, 1.0, 3.0, 1.5, 0.25)   # examples/mixture_gibbs.rs:61-65
    init = mm.init_det(64, 2)
    s = mm.GibbsSampler(mm.MixtureConditional(*params), init).set_seed(9).set_chain_offset(1000)
    trace = np.zeros((64, 120, 2))
    out = s.run(100, 20, trace=trace)
    rep = oracle.gibbs_run(oracle.G_MIXTURE2, params, init, 100, 20, tapes=(trace[:, :, 0], trace[:, :, 1]))
    np.testing.assert_array_equal(out[:, :, 1], rep["out"][:, :, 1])
    np.testing.assert_allclose(out[:, :, 0], rep["out"][:, :, 0], rtol=1e-14)
    # Philox is keyed by the global chain id and the step: shards and continuation reproduce the same draws
    a = mm.GibbsSampler(mm.MixtureConditional(*params), init[32:]).set_seed(9).set_chain_offset(1032)
    part = np.concatenate([a.run(0, 20), a.run(60, 0), a.run(40, 0)], axis=1)
    np.testing.assert_array_equal(part, out[32:])
    u, z = trace[:, :, 1], trace[:, :, 0]
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(z.mean()) < 0.05 and abs(z.std() - 1.0) < 0.05


@pytest.mark.parametrize("params", [(-2.0, 1.0, 3.0, 1.5, 0.5), (-42.0, 69.0, 1.0, 2.0, 0.123)])
def test_mixture_moments_like_reference_tests(mm, params):
    # assert_mixture_simulation, src/gibbs.rs:327-376: 4 chains x (100000 + 10000), mean and variance within 10 %
    mu0, s0, mu1, s1, pi0 = params
    theo_mean = pi0 * mu0 + (1 - pi0) * mu1
    theo_var = pi0 * (s0 ** 2 + (mu0 - theo_mean) ** 2) + (1 - pi0) * (s1 ** 2 + (mu1 - theo_mean) ** 2)
    s = mm.GibbsSampler(mm.MixtureConditional(*params), mm.init_det(4, 2)).set_seed(42)
    x = s.run(100_000, 10_000)[:, :, 0].ravel()
    assert abs(x.mean() - theo_mean) < abs(theo_mean) / 10.0
    assert abs(x.var(ddof=1) - theo_var) < abs(theo_var) / 10.0


def test_gibbs_run_progress_blocks_equal_single_run(mm):
    params = (-2.0, 1.0, 3.0, 1.5, 0.25)
    init = mm.init_det(50, 2)
    a = mm.GibbsSampler(mm.MixtureConditional(*params), init).set_seed(3)
    b = mm.GibbsSampler(mm.MixtureConditional(*params), init).set_seed(3)
    seen = []
    sample, stats = a.run_progress(128, 40, progress=lambda d, i: seen.append((d, i)), block=32)
    np.testing.assert_array_equal(sample, b.run(128, 40))
    assert seen[-1][0] == 168 and np.isfinite(seen[-1][1]["max_rhat"])
    assert 0.0 < seen[-1][1]["p_accept"] <= 1.0
