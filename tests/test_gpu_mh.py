"""Parity of the CUDA Metropolis-Hastings path (K1) against the oracle, through the C ABI."""
import math

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mm(cuda_device):
    import mini_mcmc_b200 as m

    return m


# ------------------------------------------------------------------ config C1: Gaussian2D, f64
def test_c1_minimal_mh_replay_reference_streams(mm):
    """examples/minimal_mh.rs: 4 chains x (1000 + 100), fed with the reference's own SmallRng streams
    (proposal noise incl. the D+1 draw quirk, accept uniforms from SmallRng(1 + seed + i))."""
    chains, n_collect, n_discard, D = 4, 1000, 100, 2
    noise, u = oracle.mh_cont_reference_tape(42, 42, chains, n_collect + n_discard, D)
    init = oracle.init_det(chains, D)
    tp = [0.0, 0.0, 1.0, 0.0, 0.0, 1.0]
    exp, exp_state, exp_trace = oracle.mh_cont_run_replay(oracle.T_GAUSSIAN2D, tp, 1.0, init, n_collect, n_discard,
                                                          noise, u, want_trace=True)
    mh = mm.MetropolisHastings(mm.Gaussian2D([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]]), mm.IsotropicGaussian(1.0),
                               mm.init_det(chains, D))
    trace = np.zeros((chains, n_collect + n_discard, 4))
    got = mh.run(n_collect, n_discard, replay=dict(noise=noise, u=u), trace=trace)
    assert got.shape == (chains, n_collect, D)
    # accept decisions identical, states / log-probs within 1e-12 relative (fp64 tolerance of north_star)
    np.testing.assert_array_equal(trace[..., 3], exp_trace[..., 3])
    np.testing.assert_allclose(trace[..., :3], exp_trace[..., :3], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got, exp, rtol=1e-12, atol=0)
    np.testing.assert_allclose(mh.current_state(), exp_state, rtol=1e-12, atol=0)


@pytest.mark.parametrize("kind,D", [("gauss", 2), ("iso", 3), ("iso", 1)])
def test_mh_cont_replay_random_tapes(mm, kind, D):
    rng = np.random.default_rng(7)
    chains, n_collect, n_discard = 257, 40, 13
    steps = n_collect + n_discard
    noise = rng.normal(size=(chains, steps, D))
    u = rng.random((chains, steps))
    u[0, 0] = 0.0  # ln(0) = -inf must accept
    init = rng.normal(size=(chains, D))
    if kind == "gauss":
        tgt, tp, okind = mm.Gaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), [0.0, 1.0, 4.0, 2.0, 2.0, 3.0], oracle.T_GAUSSIAN2D
    else:
        tgt, tp, okind = mm.IsotropicGaussian(1.7, dim=D), [1.7], oracle.T_ISO_GAUSSIAN
    exp, exp_state, _ = oracle.mh_cont_run_replay(okind, tp, 0.8, init, n_collect, n_discard, noise, u)
    mh = mm.MetropolisHastings(tgt, mm.IsotropicGaussian(0.8), init)
    got = mh.run(n_collect, n_discard, replay=dict(noise=noise, u=u))
    np.testing.assert_allclose(got, exp, rtol=1e-12, atol=0)
    # continuation: a second run keeps going from the stored state
    noise2 = rng.normal(size=(chains, 5, D))
    u2 = rng.random((chains, 5))
    exp2, _, _ = oracle.mh_cont_run_replay(okind, tp, 0.8, exp_state, 5, 0, noise2, u2)
    got2 = mh.run(5, 0, replay=dict(noise=noise2, u=u2))
    np.testing.assert_allclose(got2, exp2, rtol=1e-12, atol=0)


def test_mh_gaussian2d_native_distribution(mm):
    """src/metropolis_hastings.rs:338-401 bounds (mean +-0.3, cov +-0.5) on the native Philox path."""
    mh = mm.MetropolisHastings(mm.Gaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), mm.IsotropicGaussian(1.0),
                               mm.init_det(64, 2)).seed(42)
    s = mh.run(2000, 500).reshape(-1, 2)
    assert np.abs(s.mean(axis=0) - [0.0, 1.0]).max() < 0.3
    assert np.abs(np.cov(s.T) - [[4.0, 2.0], [2.0, 3.0]]).max() < 0.5


# ------------------------------------------------------------------ config C2: Poisson, integer state
@pytest.mark.parametrize("mode", [0, 1])
def test_poisson_replay_bit_exact(mm, mode):
    """Integer MH must match the oracle bit for bit under replayed flips/uniforms (both accept modes)."""
    chains, n_collect, n_discard = 1000, 300, 37  # ragged: not multiples of the warp / tile sizes
    steps = n_collect + n_discard
    # the reference's accept stream: chain i <- SmallRng(1 + seed + i); flips: SmallRng(seed) per chain
    u = np.stack([oracle.SmallRng(1 + 42 + i).f64(steps) for i in range(chains)])
    flip = np.stack([oracle.SmallRng(1000 + i).bool_half(steps) for i in range(chains)])
    init = np.zeros((chains, 1), dtype=np.uint64)
    exp, exp_state = oracle.mh_poisson_run_replay(4.0, init[:, 0], n_collect, n_discard, flip, u)
    mh = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).set_accept_mode(mode)
    got = mh.run(n_collect, n_discard, replay=dict(flip=flip, u=u))
    assert got.dtype == np.uint64 and got.shape == (chains, n_collect, 1)
    np.testing.assert_array_equal(got, exp)
    np.testing.assert_array_equal(mh.current_state()[:, 0], exp_state)


@pytest.mark.parametrize("mode", [0, 1])
def test_poisson_native_philox_bit_exact(mm, mode):
    """Native Philox keying (seed, global chain, step) has an integer twin in the oracle: bit-exact,
    including chain offsets (GPU-count invariance) and continuation across run() calls."""
    chains, n_collect, n_discard = 4099, 130, 21
    init = np.zeros((chains, 1), dtype=np.uint64)
    exp, exp_state = oracle.mh_poisson_run_philox(4.0, init[:, 0], n_collect, n_discard, seed=1234, chain_offset=77)
    mh = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(1234)
    mh.set_chain_offset(77).set_accept_mode(mode)
    got = mh.run(n_collect, n_discard)
    np.testing.assert_array_equal(got, exp)
    exp2, _ = oracle.mh_poisson_run_philox(4.0, exp_state, 33, 0, seed=1234, chain_offset=77,
                                           step_base=n_collect + n_discard)
    got2 = mh.run(33, 0)
    np.testing.assert_array_equal(got2, exp2)


def test_poisson_sharding_invariance(mm):
    """Two shards with chain offsets reproduce the single-GPU result exactly."""
    chains = 2048
    init = np.zeros((chains, 1), dtype=np.uint64)
    full = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(5).run(100, 10)
    a = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init[:1000]).seed(5).run(100, 10)
    b = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init[1000:]).seed(5)
    b = b.set_chain_offset(1000).run(100, 10)
    np.testing.assert_array_equal(np.concatenate([a, b]), full)


def test_poisson_pmf(mm):
    """tests/metrohast_poisson_test.rs:90-130: empirical pmf within 0.05 for k = 0..10."""
    init = np.zeros((4096, 1), dtype=np.uint64)
    s = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(42).run(500, 2000)
    ks = s.reshape(-1)
    for k in range(11):
        pmf = math.exp(-4.0 + k * math.log(4.0) - math.lgamma(k + 1))
        assert abs((ks == k).mean() - pmf) < 0.01


def test_poisson_large_device_resident_properties(mm):
    """Full-width launch (1,048,576 chains) with the draws left in HBM: size-independent properties —
    every transition moves by at most 1, values stay in range, the pmf is right, and a spot-check of
    rows against the oracle's Philox twin."""
    import torch

    chains, n_collect, n_discard = 1 << 20, 128, 64
    init = np.zeros((chains, 1), dtype=np.uint64)
    mh = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(99)
    out = mh.run_device(n_collect, n_discard)[:, :, 0]
    assert out.shape == (chains, n_collect)
    d = (out[:, 1:] - out[:, :-1]).abs()
    assert int(d.max()) <= 1 and int(out.min()) >= 0 and int(out.max()) < 64
    last = out[:, -1].float()
    assert abs(float(last.mean()) - 4.0) < 0.02 and abs(float(last.var()) - 4.0) < 0.05
    rows = [0, 1, 31, 32, 12345, chains - 1]
    exp, _ = oracle.mh_poisson_run_philox(4.0, np.zeros(1, dtype=np.uint64), n_collect, n_discard, seed=99,
                                          chain_offset=0)
    np.testing.assert_array_equal(out[0].cpu().numpy().astype(np.uint64), exp[0, :, 0])
    for r in rows[1:]:
        e, _ = oracle.mh_poisson_run_philox(4.0, np.zeros(1, dtype=np.uint64), n_collect, n_discard, seed=99,
                                            chain_offset=r)
        np.testing.assert_array_equal(out[r].cpu().numpy().astype(np.uint64), e[0, :, 0])


def test_poisson_host_paths_agree(mm, monkeypatch):
    """mmc_mh_run's compact device->host pipeline (u8 draws over PCIe, widened to u64 on host threads) returns the
    same array as the plain u64 copy, including ragged block sizes and odd n_collect (unaligned rows)."""
    for chains, n_collect, n_discard in ((70001, 501, 7), (4096, 1000, 0)):
        init = np.zeros((chains, 1), dtype=np.uint64)
        a = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(3).run(n_collect, n_discard)
        monkeypatch.setenv("MMC_NO_COMPACT", "1")
        b = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(3).run(n_collect, n_discard)
        monkeypatch.delenv("MMC_NO_COMPACT")
        np.testing.assert_array_equal(a, b)
    # wide state range (lambda large -> u16 compact elements)
    init = np.full((3000, 1), 400, dtype=np.uint64)
    a = mm.MetropolisHastings(mm.PoissonTarget(400.0), mm.NonnegativeProposal(), init).seed(3).run(300, 10)
    exp, _ = oracle.mh_poisson_run_philox(400.0, init[:, 0], 300, 10, seed=3)
    np.testing.assert_array_equal(a, exp)


# ---------------------------------------------------------------- Categorical target (src/distributions.rs:422-477)
@pytest.mark.parametrize("probs", [[0.2, 0.3, 0.5], [5.0, 1.0, 0.0, 2.0, 2.0], list(np.linspace(1.0, 3.0, 300))])
def test_categorical_mh_bit_exact_with_oracle(mm, probs):
    rng = np.random.default_rng(len(probs))
    K = len(probs)
    chains, nc, nd = 200, 230, 57
    valid = [k for k in range(K) if probs[k] > 0]
    init = rng.choice(valid, size=(chains, 1)).astype(np.uint64)
    # native Philox stream vs the oracle's integer twin (same octet contract as the Poisson kernel), with a shard offset
    mh = mm.MetropolisHastings(mm.Categorical(probs), mm.NonnegativeProposal(), init).seed(21).set_chain_offset(5000)
    out = mh.run(nc, nd)
    exp, exp_state = oracle.mh_categorical_run_philox(probs, init, nc, nd, seed=21, chain_offset=5000)
    np.testing.assert_array_equal(out, exp)
    np.testing.assert_array_equal(mh.current_state().reshape(-1), exp_state)
    # continuation keeps the step counter
    out2 = mh.run(40, 0)
    exp2, _ = oracle.mh_categorical_run_philox(probs, exp_state, 40, 0, seed=21, chain_offset=5000, step_base=nc + nd)
    np.testing.assert_array_equal(out2, exp2)
    # replayed flips / uniforms (what a reference run with its own RNG streams would feed)
    flip = rng.integers(0, 2, size=(chains, nc + nd)).astype(np.uint8)
    u = rng.random((chains, nc + nd))
    mh2 = mm.MetropolisHastings(mm.Categorical(probs), mm.NonnegativeProposal(), init)
    out3 = mh2.run(nc, nd, replay=dict(flip=flip, u=u))
    exp3, _ = oracle.mh_categorical_run_replay(probs, init, nc, nd, flip, u)
    np.testing.assert_array_equal(out3, exp3)
    assert out.max() < K


def test_categorical_mh_frequencies_and_guards(mm):
    probs = [0.2, 0.3, 0.5]
    mh = mm.MetropolisHastings(mm.Categorical(probs), mm.NonnegativeProposal(), np.zeros((4096, 1), dtype=np.uint64)).seed(1)
    out = mh.run(2000, 200)
    freq = np.bincount(out.ravel().astype(np.int64), minlength=3) / out.size
    np.testing.assert_allclose(freq, probs, atol=5e-3)
    cat = mm.Categorical([2.0, 3.0, 5.0])
    np.testing.assert_allclose(cat.probs, probs)                       # normalised like Categorical::new
    assert cat.logp(3) == float("-inf") and abs(cat.logp(2) - np.log(0.5)) < 1e-15
    with pytest.raises(Exception):
        mm.MetropolisHastings(cat, mm.NonnegativeProposal(), np.full((2, 1), 3, dtype=np.uint64))   # start outside the support
    with pytest.raises(Exception):
        mh.set_accept_mode(0)


def test_poisson_compact_output_equals_widened_output(mm):
    """mmc_mh_run_compact (opt-in u8 return type) returns the very draws mmc_mh_run widens to the reference's u64."""
    chains = 3000
    init = np.zeros((chains, 1), dtype=np.uint64)
    a = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(5)
    b = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.NonnegativeProposal(), init).seed(5)
    wide = a.run(120, 30)
    compact = b.run_compact(120, 30)
    assert compact.dtype == np.uint8 and compact.shape == wide.shape
    np.testing.assert_array_equal(compact.astype(np.uint64), wide)
    # continuation: both handles advanced by the same 150 steps
    np.testing.assert_array_equal(b.run_compact(10, 0).astype(np.uint64), a.run(10, 0))


# ---------------------------------------------------------------- MetropolisHastings<f32, f32, ..> and any-dimension targets
def _f32_tapes(rng, chains, steps, D):
    """Tapes of f32-representable values (the reference draws `u: f32`, `z: f32` for MetropolisHastings<f32, ..>)."""
    noise = rng.normal(size=(chains, steps, D)).astype(np.float32).astype(np.float64)
    u = rng.random((chains, steps)).astype(np.float32).astype(np.float64)
    u[u >= 1.0] = 0.5
    return noise, u


@pytest.mark.parametrize("kind,D", [("gauss", 2), ("iso", 1), ("iso", 3), ("iso", 8), ("iso", 5), ("iso", 20)])
def test_mh_f32_state_replay(mm, kind, D):
    """f32 state (src/metropolis_hastings.rs:87): proposal, log-probabilities and the log-ratio reproduce the oracle's f32
    restatement operation by operation; accept decisions are identical away from ties of r with ln(u) (the device logf
    is not the host's), and chains whose decisions agree have bit-identical states."""
    rng = np.random.default_rng(11 + D)
    chains, steps = 300, 1
    noise, u = _f32_tapes(rng, chains, steps, D)
    u[0, 0] = 0.0
    init = rng.normal(size=(chains, D)).astype(np.float32)
    if kind == "gauss":
        tgt, tp, okind = mm.Gaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), [0.0, 1.0, 4.0, 2.0, 2.0, 3.0], oracle.T_GAUSSIAN2D
    else:
        tgt, tp, okind = mm.IsotropicGaussian(1.7, dim=D), [1.7], oracle.T_ISO_GAUSSIAN
    exp, exp_state, exp_trace = oracle.mh_cont_run_replay_f32(okind, tp, 0.8, init, 1, 0, noise, u, want_trace=True)
    mh = mm.MetropolisHastings(tgt, mm.IsotropicGaussian(0.8), init)
    assert mh._np_dtype == np.float32
    trace = np.zeros((chains, steps, 4))
    got = mh.run(1, 0, replay=dict(noise=noise, u=u), trace=trace)
    assert got.dtype == np.float32
    # single transition: log-probabilities and the ratio are the same f32 operations -> equal to the last bit
    np.testing.assert_array_equal(trace[..., :3].astype(np.float32), exp_trace[..., :3])
    lnu = np.log(np.maximum(u, 1e-300))
    tie = np.abs(exp_trace[..., 2].astype(np.float64) - lnu) < 1e-5 * np.maximum(1.0, np.abs(lnu))
    same = trace[..., 3] == exp_trace[..., 3]
    assert (same | tie).all() and tie.mean() < 0.01
    ok = same[:, 0]
    np.testing.assert_array_equal(got[ok], exp[ok])
    np.testing.assert_array_equal(mh.current_state()[ok], exp_state[ok])


@pytest.mark.parametrize("D", [5, 16, 100, 256])
def test_mh_iso_any_dimension_f64(mm, D):
    """IsotropicGaussian target + proposal beyond the register-resident dimensions: sequential sums like the CPU code."""
    rng = np.random.default_rng(D)
    chains, n_collect, n_discard = 130, 12, 5
    steps = n_collect + n_discard
    noise = rng.normal(size=(chains, steps, D))
    u = rng.random((chains, steps))
    init = rng.normal(size=(chains, D))
    exp, exp_state, _ = oracle.mh_cont_run_replay(oracle.T_ISO_GAUSSIAN, [1.3], 0.3, init, n_collect, n_discard, noise, u)
    mh = mm.MetropolisHastings(mm.IsotropicGaussian(1.3, dim=D), mm.IsotropicGaussian(0.3), init)
    got = mh.run(n_collect, n_discard, replay=dict(noise=noise, u=u))
    np.testing.assert_allclose(got, exp, rtol=1e-12, atol=0)
    np.testing.assert_allclose(mh.current_state(), exp_state, rtol=1e-12, atol=0)
    with pytest.raises(Exception):
        mm.MetropolisHastings(mm.IsotropicGaussian(1.0, dim=257), mm.IsotropicGaussian(0.3), np.zeros((2, 257)))


def test_mh_f32_native_distribution_and_progress(mm):
    """Native Philox path on f32 state: the Gaussian2D moments of src/metropolis_hastings.rs:338-401, sharding invariance
    and run_progress == run."""
    init = mm.init_det(64, 2).astype(np.float32)
    tgt = mm.Gaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    s = mm.MetropolisHastings(tgt, mm.IsotropicGaussian(1.0), init).seed(42).run(2000, 500)
    assert s.dtype == np.float32
    flat = s.reshape(-1, 2).astype(np.float64)
    assert np.abs(flat.mean(axis=0) - [0.0, 1.0]).max() < 0.3
    assert np.abs(np.cov(flat.T) - [[4.0, 2.0], [2.0, 3.0]]).max() < 0.5
    part = mm.MetropolisHastings(tgt, mm.IsotropicGaussian(1.0), init[40:]).seed(42).set_chain_offset(40).run(2000, 500)
    np.testing.assert_array_equal(part, s[40:])
    sample, stats = mm.MetropolisHastings(tgt, mm.IsotropicGaussian(1.0), init).seed(42).run_progress(2000, 500, progress=False, block=300)
    np.testing.assert_array_equal(sample, s)


# ---------------------------------------------------------------- i32 variants of tests/metrohast_poisson_test.rs
@pytest.mark.parametrize("case", ["poisson", "binomial"])
def test_tabulated_reflecting_walk_bit_exact(mm, case):
    """PoissonDist + PoissonRandomWalk and BinomialDist + BinomialRandomWalk (tests/metrohast_poisson_test.rs:18-85,
    157-214): the device's table-driven kernel against the oracle's literal `r > ln(u)` step, native Philox stream and
    replayed flips / uniforms."""
    rng = np.random.default_rng(3)
    if case == "poisson":
        tgt, upper, start = mm.TabulatedTarget.poisson(4.0, 64), -1, 0
    else:
        tgt, upper, start = mm.TabulatedTarget.binomial(10, 0.3), 10, 5
    chains, nc, nd = 333, 210, 45
    init = np.full((chains, 1), start, dtype=np.uint64)
    mh = mm.MetropolisHastings(tgt, mm.ReflectingRandomWalk(), init).seed(42).set_chain_offset(9)
    out = mh.run(nc, nd)
    exp, exp_state = oracle.mh_tabulated_run_philox(tgt.table, init, nc, nd, seed=42, reflect=True, upper=upper, chain_offset=9)
    np.testing.assert_array_equal(out, exp)
    np.testing.assert_array_equal(mh.current_state().reshape(-1), exp_state)
    out2 = mh.run(30, 0)
    exp2, _ = oracle.mh_tabulated_run_philox(tgt.table, exp_state, 30, 0, seed=42, reflect=True, upper=upper, chain_offset=9,
                                             step_base=nc + nd)
    np.testing.assert_array_equal(out2, exp2)
    flip = rng.integers(0, 2, size=(chains, nc + nd)).astype(np.uint8)
    u = rng.random((chains, nc + nd))
    out3 = mm.MetropolisHastings(tgt, mm.ReflectingRandomWalk(), init).run(nc, nd, replay=dict(flip=flip, u=u))
    exp3, _ = oracle.mh_tabulated_run_replay(tgt.table, init, nc, nd, flip, u, reflect=True, upper=upper)
    np.testing.assert_array_equal(out3, exp3)
    # the same table under NonnegativeProposal follows examples/poisson_mh.rs' asymmetric walk
    out4 = mm.MetropolisHastings(tgt, mm.NonnegativeProposal(), init).run(nc, nd, replay=dict(flip=flip, u=u))
    exp4, _ = oracle.mh_tabulated_run_replay(tgt.table, init, nc, nd, flip, u, reflect=False)
    np.testing.assert_array_equal(out4, exp4)


def test_tabulated_pmf_pins(mm):
    """test_poisson_mh / test_binomial_mh (tests/metrohast_poisson_test.rs:90-130,220-249): frequencies of k = 0..10
    within 0.05 of the pmf (0.01 here: many chains)."""
    init = np.zeros((2048, 1), dtype=np.uint64)
    s = mm.MetropolisHastings(mm.PoissonTarget(4.0), mm.ReflectingRandomWalk(), init).seed(42).run(1000, 2000).reshape(-1)
    for k in range(11):
        pmf = math.exp(-4.0 + k * math.log(4.0) - math.lgamma(k + 1))
        assert abs((s == k).mean() - pmf) < 0.01
    init = np.full((2048, 1), 5, dtype=np.uint64)
    s = mm.MetropolisHastings(mm.TabulatedTarget.binomial(10, 0.3), mm.ReflectingRandomWalk(), init).seed(42).run(1000, 2000).reshape(-1)
    assert s.max() <= 10
    for k in range(11):
        pmf = math.comb(10, k) * 0.3 ** k * 0.7 ** (10 - k)
        assert abs((s == k).mean() - pmf) < 0.01
