"""Pins the oracle (CPU restatement) against every golden vector / known-answer test the reference's own
test-suite holds for the sampler hot path.  No GPU needed.

Reference locations (relative to /root/reference) are quoted next to each expected value; the values are
copied from the reference's assertions, not generated here.
"""
import math

import numpy as np
import pytest

import oracle


# ---------------------------------------------------------------- NUTS known answers
def test_find_reasonable_epsilon():
    # src/nuts.rs:1049-1055
    eps = oracle.nuts_find_reasonable_epsilon(oracle.std_normal(2), [0.0, 1.0], [1.0, 0.0])
    assert eps == 2.0


def test_build_tree():
    # src/nuts.rs:1057-1121, tolerance rel 1e-5 / abs 1e-6 as in the reference
    tgt = oracle.diff_gaussian2d([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    r = oracle.nuts_build_tree(tgt, [0.0, 1.0], [2.0, 3.0], [4.0, 5.0], logu=-2.0, v=-1, j=3, epsilon=0.01,
                               joint_0=0.1, rng_seed=0)
    kw = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r["position_minus"], [-0.1584001, 0.76208336], **kw)
    np.testing.assert_allclose(r["mom_minus"], [1.9800036, 2.9718253], **kw)
    np.testing.assert_allclose(r["grad_minus"], [-7.91236e-5, 7.9358295e-2], **kw)
    np.testing.assert_allclose(r["position_plus"], [-0.0198, 0.97025], **kw)
    np.testing.assert_allclose(r["mom_plus"], [1.98, 2.9749503], **kw)
    np.testing.assert_allclose(r["grad_plus"], [-1.250e-05, 9.925e-03], **kw)
    np.testing.assert_allclose(r["position_prime"], [-0.0198, 0.97025], **kw)
    np.testing.assert_allclose(r["grad_prime"], [-1.250e-05, 9.925e-03], **kw)
    assert r["n_prime"] == 0
    assert r["s_prime"] is True
    assert r["n_alpha_prime"] == 8
    assert abs(r["logp_prime"] - (-2.8777454)) < 1e-6
    assert abs(r["alpha_prime"] - 0.0006866617) < 1e-8


def test_chain_1():
    # src/nuts.rs:1123-1136: run(1, 0) returns the starting position
    tgt = oracle.diff_gaussian2d([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    r = oracle.nuts_run(tgt, [[0.0, 1.0]], 0.8, 1, 0, seed=41)
    np.testing.assert_allclose(r["out"].reshape(-1), [0.0, 1.0], rtol=1e-5, atol=1e-6)


CHAIN_2 = [-1.168318748474121, -0.4077277183532715, -1.8463939428329468, 0.19176559150218964,
           -1.0662782192230225, -0.3948383331298828]
CHAIN_3 = [2.653707265853882, 5.560618877410889, 2.9760334491729736, 6.325948715209961, 2.187873125076294,
           5.611990928649902, 2.1512224674224854, 5.416507720947266, 2.4165120124816895, 3.9120564460754395]


def test_chain_2():
    # src/nuts.rs:1138-1162: NUTSChain seed 42  ==  NUTS.set_seed(41) chain 0
    tgt = oracle.diff_gaussian2d([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    r = oracle.nuts_run(tgt, [[0.0, 1.0]], 0.8, 3, 3, seed=41)
    np.testing.assert_allclose(r["out"].reshape(-1), CHAIN_2, rtol=1e-5, atol=1e-6)


def test_chain_3_and_run_1():
    # src/nuts.rs:1164-1222
    tgt = oracle.diff_gaussian2d([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]])
    r = oracle.nuts_run(tgt, [[-2.0, 1.0]], 0.8, 5, 5, seed=41)
    assert r["out"].shape == (1, 5, 2)
    np.testing.assert_allclose(r["out"].reshape(-1), CHAIN_3, rtol=1e-5, atol=1e-6)


def test_nuts_replay_of_recorded_reference_stream_is_identical():
    # recording the reference's own draws and replaying them must give bit-identical chains
    tgt = oracle.diff_gaussian2d([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]])
    rec = oracle.nuts_run(tgt, [[-2.0, 1.0], [0.5, 0.5]], 0.8, 5, 5, seed=41, record=True)
    rep = oracle.nuts_run(tgt, [[-2.0, 1.0], [0.5, 0.5]], 0.8, 5, 5, tapes=rec["tapes"])
    np.testing.assert_array_equal(rec["out"], rep["out"])
    np.testing.assert_array_equal(rec["counts"], rep["counts"])
    np.testing.assert_allclose(rec["out"][0].reshape(-1), CHAIN_3, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------- stats known answers
def _tracker_rhat(d0, d1):
    t = oracle.MultiChainTracker(3, 4)
    t.step(d0)
    t.step(d1)
    return t.rhat()


def test_rhat_trackers():
    eps = np.finfo(np.float32).eps * 10.0
    # src/stats.rs:703-721
    d0 = [[0, 1, 0, 1], [1, 2, 0, 2], [0, 0, 0, 2]]
    d1 = [[1, 2, 2, 0], [1, 1, 1, 1], [0, 1, 0, 0]]
    exp = np.array([math.sqrt(2.0), 1.0801234, 0.8944273, 0.8660254], dtype=np.float32)
    assert np.abs(_tracker_rhat(d0, d1) - exp).max() < eps
    # src/stats.rs:723-737 (same data as f64, expected 0.8944271)
    exp = np.array([math.sqrt(2.0), 1.0801234, 0.8944271, 0.8660254], dtype=np.float32)
    assert np.abs(_tracker_rhat(d0, d1) - exp).max() < eps
    # src/stats.rs:739-752
    d0 = [[1, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 2]]
    d1 = [[1, 2, 0, 2], [1, 2, 0, 0], [2, 0, 1, 2]]
    exp = np.array([1.0 / math.sqrt(2.0), 0.74535599, 1.0, 1.5], dtype=np.float32)
    assert np.abs(_tracker_rhat(d0, d1) - exp).max() < eps


@pytest.mark.parametrize("fn", [oracle.autocov_bf, oracle.autocov_fft])
def test_autocov_kats(fn):
    # src/stats.rs:777-790
    data = np.array([[1.0], [2.0], [3.0], [4.0]], dtype=np.float32)
    np.testing.assert_allclose(fn(data), [[1.25], [0.3125], [-0.375], [-0.5625]], atol=1e-6, rtol=0)
    # src/stats.rs:792-808
    data = np.array([[1.0, 0.3], [2.0, 2.0], [3.0, -2.0], [4.0, 5.0]], dtype=np.float32)
    exp = [[1.25, 6.516875], [0.3125, -3.7889063], [-0.375, 1.4721875], [-0.5625, -0.94171875]]
    np.testing.assert_allclose(fn(data), exp, atol=1e-6, rtol=0)


def test_ess_1():
    # src/stats.rs:810-834: 4 x 1000 f32 uniforms from SmallRng(42)
    rng = oracle.SmallRng(42)
    data = rng.f32(4000).reshape(4, 1000, 1)
    st = oracle.run_stats(data)
    assert st["ess"]["min"] > 3800.0
    assert st["rhat"]["max"] < 1.01
    # the survey's independent restatement got ESS = 4110.47, split-Rhat = 1.000116
    assert abs(st["ess"]["min"] - 4110.47) < 1.0
    assert abs(st["rhat"]["max"] - 1.000116) < 1e-5


# ---------------------------------------------------------------- distributions known answers
def _normalize_isogauss(x, d, std):
    # src/distributions.rs:567-570
    return math.exp(x - (d / 2.0) * (math.log(2.0) + math.log(math.pi) + 2.0 * math.log(std)))


def test_iso_gauss_kats():
    # src/distributions.rs:572-606
    assert abs(_normalize_isogauss(oracle.iso_unnorm_logp(1.0, [1.0]), 1, 1.0) - 0.24197072451914337) < 1e-7
    assert abs(_normalize_isogauss(oracle.iso_unnorm_logp(2.0, [0.42, 9.6]), 2, 2.0) - 3.864661987252467e-7) < 1e-15
    assert abs(_normalize_isogauss(oracle.iso_unnorm_logp(3.0, [1.0, 2.0, 3.0]), 3, 3.0)
               - 0.001080393185560214) < 1e-8


def test_gaussian2d_logp_kat():
    # src/distributions.rs:812-831
    lp = oracle.gaussian2d_logp([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]], [0.5, -0.5], normalized=True)
    assert abs(lp - (-2.0878770664093453)) < 1e-10


def test_init_det_regression():
    # core.rs:394-435; regression vector recorded in SURVEY.md §8c(iii) (restatement-derived)
    exp = [[0.8343975468437959, -0.514962928147295], [1.40772757311975, 0.46445486122523566],
           [0.9536668702127304, 0.27411555634974205], [-1.3773172567668162, 0.4144533898735936]]
    np.testing.assert_allclose(oracle.init_det(4, 2), exp, rtol=0, atol=1e-15)


# ---------------------------------------------------------------- analytic gradients vs finite differences
@pytest.mark.parametrize("tgt,x", [
    (oracle.rosenbrock_nd(5), [0.3, -0.2, 0.5, 0.1, -0.4]),
    (oracle.rosenbrock_2d(1.0, 100.0), [0.3, -0.2]),
    (oracle.diff_gaussian2d([1.0, 2.0], [[1.0, 2.0], [2.0, 5.0]]), [0.3, -0.2]),
    (oracle.std_normal(3), [0.3, -0.2, 1.5]),
])
def test_gradients_match_finite_differences(tgt, x):
    x = np.asarray(x, dtype=np.float64)
    _, g = oracle.logp_grad(tgt, x)
    h = 1e-2
    fd = np.zeros_like(x)
    for i in range(len(x)):
        xp, xm = x.copy(), x.copy()
        xp[i] += h
        xm[i] -= h
        # central differences with Richardson extrapolation keep the f32 evaluation noise small
        f = lambda z: oracle.logp_grad(tgt, z)[0]
        xp2, xm2 = x.copy(), x.copy()
        xp2[i] += 2 * h
        xm2[i] -= 2 * h
        fd[i] = (8 * (f(xp) - f(xm)) - (f(xp2) - f(xm2))) / (12 * h)
    np.testing.assert_allclose(g, fd, rtol=2e-2, atol=2e-2)


def test_rosenbrock_nd_matches_torch_autograd():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)
    for D in (2, 3, 7, 100):
        x = rng.normal(size=D).astype(np.float32)
        xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        low, high = xt[:-1], xt[1:]
        lp = -(100.0 * (high - low ** 2) ** 2 + (1 - low) ** 2).sum()
        lp.backward()
        olp, og = oracle.logp_grad(oracle.rosenbrock_nd(D), x)
        assert abs(olp - lp.item()) <= 2e-5 * max(1.0, abs(lp.item()))
        np.testing.assert_allclose(og, xt.grad.numpy(), rtol=2e-5, atol=2e-4)


# ---------------------------------------------------------------- statistical pins for MH (reference's own bounds)
def test_mh_gaussian2d_statistical_pin():
    # src/metropolis_hastings.rs:338-401 (3 chains, 500 burn-in, seed 42, mean +-0.3 / cov +-0.5)
    chains, n_collect, burn = 3, 5000, 500
    noise, u = oracle.mh_cont_reference_tape(42, 42, chains, n_collect + burn, 2)
    init = oracle.init_det(chains, 2)
    tp = [0.0, 1.0, 4.0, 2.0, 2.0, 3.0]
    out, _, _ = oracle.mh_cont_run_replay(oracle.T_GAUSSIAN2D, tp, 1.0, init, n_collect, burn, noise, u)
    flat = out.reshape(-1, 2)
    mean = flat.mean(axis=0)
    cov = np.cov(flat.T)
    assert np.abs(mean - [0.0, 1.0]).max() < 0.3
    assert np.abs(cov - [[4.0, 2.0], [2.0, 3.0]]).max() < 0.5


def test_mh_poisson_statistical_pin():
    # tests/metrohast_poisson_test.rs:90-130: pmf within 0.05 for k = 0..10, 20k draws after 2k burn-in
    out, _ = oracle.mh_poisson_run_reference(4.0, np.zeros(1, dtype=np.uint64), 20000, 2000, seed=42)
    ks = out.reshape(-1)
    for k in range(11):
        pmf = math.exp(-4.0 + k * math.log(4.0) - math.lgamma(k + 1))
        assert abs((ks == k).mean() - pmf) < 0.05


def test_poisson_kernels_agree():
    # replay and philox entry points share one transition function
    rng = np.random.default_rng(1)
    chains, steps = 8, 200
    flip = rng.integers(0, 2, size=(chains, steps), dtype=np.uint8)
    u = rng.random((chains, steps))
    out, st = oracle.mh_poisson_run_replay(4.0, np.zeros(chains, dtype=np.uint64), steps, 0, flip, u)
    assert out.shape == (chains, steps, 1) and (out[:, -1, 0] == st).all()
    # hand-rolled transition for chain 0
    x = 0
    for i in range(steps):
        y = 1 if x == 0 else (x + 1 if flip[0, i] else x - 1)
        r = (oracle.poisson_logp(4.0, y) + oracle.nonneg_logq(y, x)) - (oracle.poisson_logp(4.0, x) + oracle.nonneg_logq(x, y))
        if r > (math.log(u[0, i]) if u[0, i] > 0 else -math.inf):
            x = y
        assert out[0, i, 0] == x


# ---------------------------------------------------------------- Gibbs (src/gibbs.rs tests)
@pytest.mark.parametrize("params", [(-2.0, 1.0, 3.0, 1.5, 0.5), (-42.0, 69.0, 1.0, 2.0, 0.123)])
def test_gibbs_mixture_statistical_pins(params):
    # assert_mixture_simulation(…, 4 chains, 100_000 + 10_000, seed 42), src/gibbs.rs:327-410
    mu0, s0, mu1, s1, pi0 = params
    init = oracle.init_det(4, 2) if hasattr(oracle, "init_det") else np.zeros((4, 2))
    r = oracle.gibbs_run(oracle.G_MIXTURE2, params, init, 100_000, 10_000, cond_seed=42)
    x = r["out"][:, :, 0].ravel()
    theo_mean = pi0 * mu0 + (1 - pi0) * mu1
    theo_var = pi0 * (s0 ** 2 + (mu0 - theo_mean) ** 2) + (1 - pi0) * (s1 ** 2 + (mu1 - theo_mean) ** 2)
    assert abs(x.mean() - theo_mean) < abs(theo_mean) / 10.0
    assert abs(x.var(ddof=1) - theo_var) < abs(theo_var) / 10.0


def test_gibbs_constant_and_replay_roundtrip():
    r = oracle.gibbs_run(oracle.G_CONSTANT, [42.0], np.zeros((4, 2)), 10, 5)   # src/gibbs.rs:291-305
    np.testing.assert_array_equal(r["out"], np.full((4, 10, 2), 42.0))
    init = np.array([[0.3, 0.0], [-1.0, 1.0]])
    rec = oracle.gibbs_run(oracle.G_MIXTURE2, (-2.0, 1.0, 3.0, 1.5, 0.25), init, 50, 10, cond_seed=7, record=True)
    rep = oracle.gibbs_run(oracle.G_MIXTURE2, (-2.0, 1.0, 3.0, 1.5, 0.25), init, 50, 10, tapes=rec["tapes"])
    np.testing.assert_array_equal(rec["out"], rep["out"])


def test_tabulated_reflecting_walk_pmf_pins():
    """test_poisson_mh / test_binomial_mh (tests/metrohast_poisson_test.rs:90-130,220-249): one chain, 20,000 + 2,000
    transitions from 0 (Poisson) / 5 (Binomial), frequencies of k = 0..10 within 0.05 of the pmf - on the oracle's
    restatement of the i32 targets with the reflecting +-1 walk."""
    import math

    import mini_mcmc_b200.distributions as dist

    for table, start, upper, pmf in (
        (dist.TabulatedTarget.poisson(4.0, 64).table, 0, -1, lambda k: math.exp(-4.0 + k * math.log(4.0) - math.lgamma(k + 1))),
        (dist.TabulatedTarget.binomial(10, 0.3).table, 5, 10, lambda k: math.comb(10, k) * 0.3 ** k * 0.7 ** (10 - k)),
    ):
        flips = oracle.SmallRng(42).bool_half(22_000)[None]
        u = oracle.SmallRng(43).f64(22_000)[None]
        out, _ = oracle.mh_tabulated_run_replay(table, np.array([start], dtype=np.uint64), 20_000, 2_000, flips, u,
                                                reflect=True, upper=upper)
        s = out.reshape(-1)
        assert s.max() <= (upper if upper >= 0 else 63)
        for k in range(11):
            assert abs((s == k).mean() - pmf(k)) < 0.05
    # BinomialDist's table is ln C(n, k) + k ln p + (n - k) ln(1 - p)
    t = dist.TabulatedTarget.binomial(10, 0.3).table
    np.testing.assert_allclose(np.exp(t), [math.comb(10, k) * 0.3 ** k * 0.7 ** (10 - k) for k in range(11)], rtol=1e-12)
