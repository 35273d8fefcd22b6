"""Device-side progress trackers (ChainTracker / MultiChainTracker / collect_rhat, src/stats.rs:26-307) and the
block-wise run_progress of the three samplers, against the oracle restatement and the reference's own known answers."""
import math

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mm(cuda_device):
    import mini_mcmc_b200 as m

    return m


# ---------------------------------------------------------------- MultiChainTracker
def test_multichain_tracker_reference_known_answers(mm):
    eps = np.finfo(np.float32).eps * 10.0
    cases = [
        # src/stats.rs:703-721 (f32) and :723-737 (f64 input, expected 0.8944271)
        ([[0, 1, 0, 1], [1, 2, 0, 2], [0, 0, 0, 2]], [[1, 2, 2, 0], [1, 1, 1, 1], [0, 1, 0, 0]],
         [math.sqrt(2.0), 1.0801234, 0.8944273, 0.8660254]),
        # src/stats.rs:739-752
        ([[1, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 2]], [[1, 2, 0, 2], [1, 2, 0, 0], [2, 0, 1, 2]],
         [1.0 / math.sqrt(2.0), 0.74535599, 1.0, 1.5]),
    ]
    for d0, d1, exp in cases:
        for dt in (np.float32, np.float64):
            t = mm.MultiChainTracker(3, 4)
            t.step(np.array(d0, dtype=dt))
            t.step(np.array(d1, dtype=dt))
            assert np.abs(t.rhat() - np.array(exp, dtype=np.float32)).max() < eps


def _sticky_walk(rng, c, n, p, dtype):
    """draws with repeated rows (rejections) so that the accept EMA sees both outcomes"""
    x = np.zeros((c, n, p))
    cur = rng.normal(size=(c, p))
    for t in range(n):
        move = rng.random(c) < 0.6
        cur = np.where(move[:, None], cur + rng.normal(size=(c, p)), cur)
        x[:, t] = cur
    if dtype == np.uint64:
        return np.abs(np.round(x * 3)).astype(np.uint64)
    return x.astype(dtype)


@pytest.mark.parametrize("c,n,p,dtype", [
    (4, 50, 1, np.uint64), (5, 37, 2, np.float64), (7, 64, 3, np.float32), (33, 20, 8, np.float32),
    (3, 45, 9, np.float32), (6, 70, 40, np.float64), (4, 33, 100, np.float32),
])
def test_multichain_tracker_matches_oracle(mm, c, n, p, dtype):
    rng = np.random.default_rng(c * 100 + p)
    x = _sticky_walk(rng, c, n, p, dtype)
    ref = oracle.MultiChainTracker(c, p)
    for t in range(n):
        ref.step(x[:, t].astype(np.float32))
    dev = mm.MultiChainTracker(c, p)
    dev.steps(x, 0, 11).steps(x, 11, 1).steps(x, 12)   # blocks of any size reproduce the step-by-step fold
    mean, msq, pa = dev.raw()
    np.testing.assert_array_equal(mean, ref.mean)       # same f32 recurrences: bit-exact
    np.testing.assert_array_equal(msq, ref.mean_sq)
    np.testing.assert_allclose(pa[0], ref.p_accept, rtol=1e-6)
    s = dev.summary()
    assert s["n"] == n
    np.testing.assert_allclose(s["rhat"], ref.rhat(), rtol=1e-5)
    np.testing.assert_allclose(s["max_rhat"], ref.rhat().max(), rtol=1e-5)
    np.testing.assert_allclose(s["p_accept"], ref.p_accept, rtol=1e-6)


def test_multichain_accept_window_with_many_chains(mm):
    # more chains than the EMA window: only the tail of the last step can matter
    rng = np.random.default_rng(5)
    c, n, p = 40000, 3, 2
    x = _sticky_walk(rng, c, n, p, np.float32)
    ref = oracle.MultiChainTracker(c, p)
    for t in range(n):
        ref.step(x[:, t])
    dev = mm.MultiChainTracker(c, p).steps(x)
    np.testing.assert_allclose(dev.summary()["p_accept"], ref.p_accept, rtol=1e-6)
    np.testing.assert_allclose(dev.summary()["rhat"], ref.rhat(), rtol=2e-4)   # reference sums 40000 f32 means sequentially


# ---------------------------------------------------------------- ChainTracker + collect_rhat
@pytest.mark.parametrize("c,n,p,dtype", [(4, 60, 1, np.uint64), (6, 41, 3, np.float64), (5, 50, 12, np.float32), (40, 9, 2, np.float32)])
def test_chain_trackers_match_oracle(mm, c, n, p, dtype):
    rng = np.random.default_rng(c * 7 + p)
    x = _sticky_walk(rng, c, n + 1, p, dtype)
    init, x = x[:, 0], x[:, 1:]
    refs = [oracle.ChainTracker(p, init[i]) for i in range(c)]
    for t in range(n):
        for i in range(c):
            refs[i].step(x[i, t])
    dev = mm.ChainTrackers(p, init)
    dev.steps(x, 0, 7).steps(x, 7)
    mean, msq, pa = dev.raw()
    np.testing.assert_array_equal(mean, np.stack([r.mean for r in refs]))
    np.testing.assert_array_equal(msq, np.stack([r.mean_sq for r in refs]))
    np.testing.assert_array_equal(pa, np.array([r.p_accept for r in refs], dtype=np.float32))
    exp = oracle.collect_rhat([r.stats() for r in refs])
    s = dev.summary()
    np.testing.assert_allclose(s["rhat"], exp, rtol=1e-5)
    np.testing.assert_allclose(s["p_accept"], np.mean([r.p_accept for r in refs]), rtol=1e-6)


# ---------------------------------------------------------------- run_progress in blocks == one launch
def test_hmc_run_progress_blocks_equal_single_run(mm):
    init = mm.init_with_seed(64, 3, 42, dtype=np.float32)
    a = mm.HMC(mm.RosenbrockND(), init, 0.01, 10).set_seed(3)
    b = mm.HMC(mm.RosenbrockND(), init, 0.01, 10).set_seed(3)
    seen = []
    sample, stats = a.run_progress(100, 20, progress=lambda done, info: seen.append((done, info)), block=32)
    plain = b.run_device(100, 20)
    np.testing.assert_array_equal(sample.cpu().numpy(), plain.cpu().numpy())
    assert [d for d, _ in seen] == [32, 64, 96, 100]
    # the tracker saw the post-burn-in start plus the 100 draws, like src/hmc.rs:242-266
    ref = oracle.MultiChainTracker(64, 3)
    c = mm.HMC(mm.RosenbrockND(), init, 0.01, 10).set_seed(3)
    c.run_device(0, 20)
    ref.step(c.positions)
    for t in range(100):
        ref.step(plain[:, t].cpu().numpy())
    np.testing.assert_allclose(seen[-1][1]["max_rhat"], ref.rhat().max(), rtol=1e-5)
    np.testing.assert_allclose(seen[-1][1]["p_accept"], ref.p_accept, rtol=1e-6)
    assert np.isfinite(stats.ess.min) and stats.rhat.max > 0


def test_mh_run_progress_blocks_equal_single_run(mm):
    init = np.zeros((96, 1), dtype=np.uint64)
    a = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(11)
    b = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(11)
    seen = []
    sample, stats = a.run_progress(200, 70, progress=lambda done, info: seen.append((done, info)), block=64)
    plain = b.run(200, 70)
    np.testing.assert_array_equal(sample, plain)
    assert seen[-1][0] == 270
    # one ChainTracker per chain over ALL steps (src/core.rs:90-136)
    c = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(11)
    full = c.run(270, 0)
    refs = [oracle.ChainTracker(1, init[i]) for i in range(96)]
    for t in range(270):
        for i in range(96):
            refs[i].step(full[i, t])
    np.testing.assert_allclose(seen[-1][1]["rhat"], oracle.collect_rhat([r.stats() for r in refs]), rtol=1e-5)
    np.testing.assert_allclose(seen[-1][1]["p_accept"], np.mean([r.p_accept for r in refs]), rtol=1e-6)
    # continuous target: f64 draws
    g = mm.MetropolisHastings(mm.Gaussian2D([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]]), mm.IsotropicGaussian(1.0), mm.init_det(4, 2)).seed(42)
    h = mm.MetropolisHastings(mm.Gaussian2D([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]]), mm.IsotropicGaussian(1.0), mm.init_det(4, 2)).seed(42)
    s2, _ = g.run_progress(96, 32, progress=False, block=32)
    np.testing.assert_array_equal(s2, h.run(96, 32))


def test_nuts_run_progress_blocks_equal_single_run(mm):
    init = mm.init_with_seed(40, 10, 42, dtype=np.float32)
    a = mm.NUTS(mm.RosenbrockND(), init, 0.9, scalar_dtype="f64").set_seed(5)
    b = mm.NUTS(mm.RosenbrockND(), init, 0.9, scalar_dtype="f64").set_seed(5)
    seen = []
    sample, _ = a.run_progress(64, 64, progress=lambda done, info: seen.append((done, info)), block=32)
    plain, _ = b.run_progress(64, 64)
    np.testing.assert_array_equal(sample.cpu().numpy(), plain.cpu().numpy())
    np.testing.assert_array_equal(a.state(), b.state())
    assert [d for d, _ in seen] == [32, 64, 96, 128]
    assert 0.0 < seen[-1][1]["p_accept"] <= 1.0 and np.isfinite(seen[-1][1]["max_rhat"])
